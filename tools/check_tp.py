"""row-sharded decode on N GPUs against the UNSHARDED model and against the oracle restatement.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_tp.py [layers]

1. Llama-3.1-8B shapes, `layers` layers, same seed everywhere:  world = 1 (rank-local, un-fused launch list) is the reference;
   row-sharded with the NVLink peer exchange (p2p) and with ncclAllGather, eager and graph-replayed, must produce the same
   tokens and logits within rel-L2 1e-3 (fp32 atomics order differs run to run).
2. a small model (hidden 512) whose weights the numpy oracle decodes in seconds: the row-sharded logits against the float64
   restatement with the graph's fp16 rounding points (tests/_restate.py), rel-L2 <= 1e-3.
"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "q-palette_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p_)
local, rank, world = int(os.environ["LOCAL_RANK"]), int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from qpalette.decode import DecodeRunner, LLAMA31_8B, LlamaShape, uniform_qdict
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 3
TOL = 2e-3  # whole-model logits: 2x the per-layer tolerance (run-to-run noise of the same launch list is 4e-4 .. 8e-4)


def run(shape, qd, mi, mode, nl, steps=4, graph=False, fused=False, ref_fn=None):
    kw = dict(max_seq=64, seed=5, num_layers=nl)
    if mode == "single":
        r = DecodeRunner(shape, qd, mi, fused=fused, **kw)
    else:  # p2p: NVLink peer exchange + fused GEMV prologues; p2p_unfused: peer exchange inside the glue kernels; nccl: baseline
        r = DecodeRunner(shape, qd, mi, rank=rank, world=world, process_group=dist.group.WORLD, p2p=mode.startswith("p2p"),
                         fused=(mode == "p2p"), **kw)
        assert r.fused == (mode == "p2p"), (mode, r.fused, r.p2p)
    r.reset(3)
    if graph:
        r.capture()
        r.reset(3)
    toks, logits = [], []
    for _ in range(steps):
        r.step()
        torch.cuda.synchronize()
        toks.append(int(r.token.item()))
        logits.append(r.logits.float().clone())
    t_us = None
    if graph:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        a.record()
        for _ in range(20):
            r.step()
        b.record(); torch.cuda.synchronize()
        t_us = a.elapsed_time(b) * 1e3 / 20
    extra = ref_fn(r) if ref_fn else None
    return toks, logits, t_us, extra


ok = True
shape = LLAMA31_8B
qd, mi = uniform_qdict(shape, "tcomb_6_7_0.5_none_0.9"), [["merge_qkv", "merge_ug"]] * 32
ref_t, ref_l, _, _ = run(shape, qd, mi, "single", layers)
for mode, graph, fused in (("single", True, True), ("p2p", False, False), ("p2p", True, False), ("p2p_unfused", False, False),
                           ("p2p_unfused", True, False), ("nccl", False, False), ("nccl", True, False)):
    t, l, us, _ = run(shape, qd, mi, mode, layers, graph=graph, fused=fused)
    err = max(float((a - b).norm() / b.norm()) for a, b in zip(l, ref_l))
    good = t == ref_t and err <= TOL
    ok &= good
    print(f"[rank {rank}] 8B x{layers} {mode:11s} graph={graph}: tokens {t} vs single {ref_t}  max rel-L2 of logits vs single "
          f"{err:.2e}  {'' if us is None else f'{us:.1f} us/step'}  {'OK' if good else 'MISMATCH'}", flush=True)

# ---- small model against the oracle restatement ------------------------------------------------------------------------
import _restate as R
small = LlamaShape(hidden_size=512, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=8,
                   num_key_value_heads=2 * world if world > 2 else 2, vocab_size=1024)
if small.num_attention_heads % small.num_key_value_heads or small.num_attention_heads % world:
    small = LlamaShape(hidden_size=1024, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=16,
                       num_key_value_heads=world, vocab_size=1024)
sqd, smi = uniform_qdict(small, "tcomb_6_7_0.5_none_0.9"), [["merge_qkv", "merge_ug"]] * 2
single = DecodeRunner(small, sqd, smi, max_seq=64, seed=5, fused=False)   # same seed -> same full-width weights: the oracle side
caches = [([], []) for _ in single.layers]
tok, refs = 3, []
for step in range(3):  # the oracle side follows its OWN greedy tokens, independently of any GPU run
    refs.append(R.decode_step_ref(single, single.embed[tok].cpu().numpy(), step, caches)[1])
    tok = int(np.argmax(refs[step]))
for mode in ("p2p", "p2p_unfused", "nccl"):
    r = DecodeRunner(small, sqd, smi, max_seq=64, seed=5, rank=rank, world=world, process_group=dist.group.WORLD,
                     p2p=mode.startswith("p2p"), fused=(mode == "p2p"))
    r.reset(3)
    for step in range(3):
        r.step()
        torch.cuda.synchronize()
        lg = r.logits.float().cpu().numpy().astype(np.float64)
        err = float(np.linalg.norm(lg - refs[step]) / np.linalg.norm(refs[step]))
        good = err <= TOL and int(r.token.item()) == int(np.argmax(refs[step]))
        ok &= good
        print(f"[rank {rank}] small {mode:11s} step {step}: rel-L2 of logits vs oracle restatement {err:.2e} "
              f"{'OK' if good else 'MISMATCH'}", flush=True)
dist.barrier(); torch.cuda.synchronize()
print(f"[rank {rank}] {'ALL OK' if ok else 'FAILED'}", flush=True)
sys.stdout.flush()
os._exit(0 if ok else 1)
