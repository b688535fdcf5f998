"""N > 1 host logic on CPU: row shards of packed layers under a world_size-2 gloo group.  Each rank decodes only its shard
with the oracle, computes its partial GEMV, and the all-gathered result must equal the single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    for p in (ROOT, os.path.join(ROOT, "q-palette_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import qp_oracle as O
    from qpalette.shard import gather_rows, shard_plan, shard_rows, unshard_index
    from qpalette.utils.mem_op import get_dummy_quant_results
    torch.manual_seed(0)  # same synthetic layer on every rank
    K, sizes = 256, [256, 64, 64]  # merge_qkv-like: q, k, v members
    M = sum(sizes)
    ok = True
    for qs in ("tcq_6_none_0.9", "tcomb_6_7_0.5_none_0.9", "ldlq_2_8_none_1.0", "ldlq_1_4_none_1.0"):
        info = get_dummy_quant_results(None, None, qs, device="cpu", in_features=K, out_features=M)["linear_info"]
        x = torch.randn(1, K).half().numpy()
        plan = shard_plan(sizes, rank, world)
        m_loc = sum(n for _, n in plan)

        def dec(get, m):
            if "trellis" in info:
                return O.tcq_decode(get(info["trellis"]).numpy(), info["tlut"].numpy(), m, K, info["KV"], info["tlut_bits"])
            if "trellis1" in info:
                return O.tcq_decode_combt(get(info["trellis1"]).numpy(), get(info["trellis2"]).numpy(), info["tlut"].numpy(),
                                          m, K, info["KV"][0], info["KV"][1], info["tlut_bits"])
            return O.lut_tc_decode(get(info["qweight"]).numpy(), info["lut"].numpy(), m, K, info["lut_bits"], info["vec_sz"])

        W_full = dec(lambda t: t.reshape(-1), M)
        W_loc = dec(lambda t: shard_rows(t, M, plan), m_loc)
        rows = np.concatenate([np.arange(r0, r0 + n) for r0, n in plan])
        ok &= np.array_equal(W_loc.view(np.uint16), W_full[rows].view(np.uint16))  # shard == those rows, bit-exact
        y_loc = torch.from_numpy(O.gemv_ref(W_loc, x).reshape(-1))
        gathered = gather_rows(y_loc, world)
        full = torch.empty(M, dtype=gathered.dtype)
        full[unshard_index(sizes, world)] = gathered
        ok &= np.allclose(full.numpy(), O.gemv_ref(W_full, x).reshape(-1), rtol=1e-12, atol=1e-12)
    results[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharding_world2_gloo():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def test_shard_plan_rules():
    sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
    from qpalette.shard import shard_plan, unshard_index
    assert shard_plan([4096, 1024, 1024], 1, 2) == [(2048, 2048), (4096 + 512, 512), (5120 + 512, 512)]
    with pytest.raises(ValueError):
        shard_plan([1024], 0, 64)  # 16 rows per rank: not whole 32-row strips
    idx = unshard_index([128, 64], 2)
    assert sorted(idx.tolist()) == list(range(192))
