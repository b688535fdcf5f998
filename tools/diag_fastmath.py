"""what --use_fast_math changes: run with QP_LIB_SUFFIX=<variant> to dump the logits of a 4-layer Llama-3.1-8B-shaped model and
of the small oracle-checked model; run with `compare a b` to print the rel-L2 between two dumps and each one's error against the
float64 restatement (tests/_restate.py)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "q-palette_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p_)
OUT = os.path.join(ROOT, "gpurun_out")
if len(sys.argv) > 1 and sys.argv[1] == "compare":
    a, b = (torch.load(os.path.join(OUT, f"fm_logits{s}.pt"), weights_only=False) for s in sys.argv[2:4])
    rel = lambda x, y: float(np.linalg.norm(x - y) / np.linalg.norm(y))
    for k in a:
        if k.startswith("ref"):
            continue
        print(f"{k}: rel-L2 {sys.argv[2]} vs {sys.argv[3]} = {rel(a[k], b[k]):.3e}", end="")
        if "ref_" + k in a:
            print(f"   vs restatement: {sys.argv[2]} {rel(a[k], a['ref_' + k]):.3e}  {sys.argv[3]} {rel(b[k], a['ref_' + k]):.3e}", end="")
        print()
    sys.exit(0)
from qpalette.decode import DecodeRunner, LLAMA31_8B, LlamaShape, uniform_qdict
import _restate as R
res = {}
qs = "tcomb_6_7_0.5_none_0.9"
r = DecodeRunner(LLAMA31_8B, uniform_qdict(LLAMA31_8B, qs), [["merge_qkv", "merge_ug"]] * 32, max_seq=64, seed=5, num_layers=4, fused=False)
r.reset(3)
for s in range(4):
    r.step(); torch.cuda.synchronize()
    res[f"8B x4 step {s}"] = r.logits.float().cpu().numpy().astype(np.float64)
small = LlamaShape(hidden_size=512, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=2, vocab_size=1024)
r = DecodeRunner(small, uniform_qdict(small, qs), [["merge_qkv", "merge_ug"]] * 2, max_seq=64, seed=5, fused=False)
caches = [([], []) for _ in r.layers]
r.reset(3)
tok = 3
for s in range(3):
    ref = R.decode_step_ref(r, r.embed[tok].cpu().numpy(), s, caches)[1]
    r.step(); torch.cuda.synchronize()
    res[f"small step {s}"] = r.logits.float().cpu().numpy().astype(np.float64)
    res[f"ref_small step {s}"] = ref
    tok = int(np.argmax(ref))
    r.token.fill_(tok)
torch.save(res, os.path.join(OUT, f"fm_logits{os.environ.get('QP_LIB_SUFFIX', '')}.pt"))
print("saved", len(res))
