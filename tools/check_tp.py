"""row-sharded decode on N GPUs: NVLink peer exchange (p2p) vs ncclAllGather, same seed, eager and graph-replayed.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_tp.py [layers]"""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
local, rank, world = int(os.environ["LOCAL_RANK"]), int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from qpalette.decode import DecodeRunner, LLAMA31_8B, uniform_qdict
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shape = LLAMA31_8B
qd, mi = uniform_qdict(shape, "tcomb_6_7_0.5_none_0.9"), [["merge_qkv", "merge_ug"]] * 32


def run(mode, steps=4, graph=False):
    kw = dict(max_seq=64, seed=5, num_layers=layers)
    if mode == "single":
        r = DecodeRunner(shape, qd, mi, fused=False, **kw)
    else:
        r = DecodeRunner(shape, qd, mi, rank=rank, world=world, process_group=dist.group.WORLD, p2p=(mode == "p2p"), **kw)
    r.reset(3)
    if graph:
        r.capture()
        r.reset(3)
    toks, logits = [], []
    for _ in range(steps):
        r.step() if graph else r._step()
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
    return toks, logits, t_us


ok = True
ref_t, ref_l, _ = run("nccl")
for mode, graph in (("p2p", False), ("p2p", True), ("nccl", True)):
    t, l, us = run(mode, graph=graph)
    err = max(float((a - b).norm() / b.norm()) for a, b in zip(l, ref_l))
    good = t == ref_t and err < 5e-3  # fp32 atomics order differs run to run
    ok &= good
    print(f"[rank {rank}] {mode:5s} graph={graph}: tokens {t} vs {ref_t}  max rel-L2 of logits {err:.2e}  "
          f"{'' if us is None else f'{us:.1f} us/step'}  {'OK' if good else 'MISMATCH'}", flush=True)
dist.barrier(); torch.cuda.synchronize()
print(f"[rank {rank}] {'ALL OK' if ok else 'FAILED'}", flush=True)
sys.stdout.flush()
os._exit(0 if ok else 1)
