"""debug: is a fused / un-fused mismatch a race (programmatic dependent launch overlap) or deterministic?
Runs the same model four ways -- {fused, unfused} x {normal, device-synchronised after every launch} -- twice each and prints
the pairwise rel-L2 of the logits.   python tools/diag_race.py [layers=2] [config=uniform]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "q-palette_b200")):
    sys.path.insert(0, p_)
import qpalette.decode as D
from qpalette.decode import DecodeRunner, LLAMA31_8B, uniform_qdict

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name = sys.argv[2] if len(sys.argv) > 2 else "uniform"
if name == "uniform":
    qd, mi = uniform_qdict(LLAMA31_8B, "tcomb_6_7_0.5_none_0.9"), [["merge_qkv", "merge_ug"]] * 32
else:
    cfg = json.load(open(os.path.join(ROOT, "configs", name + ".json")))
    qd, mi = {k: tuple(v) for k, v in cfg["qdict"].items()}, cfg["merge_info"]
orig_check = D.check


def sync_check(rc):
    orig_check(rc)
    torch.cuda.synchronize()


def run(fused, sync, steps=2):
    D.check = sync_check if sync else orig_check
    r = DecodeRunner(LLAMA31_8B, qd, mi, max_seq=16, seed=11, num_layers=layers, fused=fused)
    r.reset(9)
    out = []
    for _ in range(steps):
        r.step(); torch.cuda.synchronize()
        out.append(r.logits.clone())
    D.check = orig_check
    del r
    torch.cuda.empty_cache()
    return out


rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
res = {}
for fused in (True, False):
    for sync in (True, False):
        if os.environ.get("QP_DIAG_QUICK") and (fused, sync) not in ((False, True), (True, False)):
            continue
        for rep in (0, 1):
            res[(fused, sync, rep)] = run(fused, sync)
ref = res[(False, True, 0)]
for k, v in res.items():
    print(f"fused={k[0]!s:5} sync={k[1]!s:5} rep={k[2]}: vs unfused+sync rep0: " + "  ".join(f"step{i} {rel(a, b):.2e}" for i, (a, b) in enumerate(zip(v, ref))), flush=True)
