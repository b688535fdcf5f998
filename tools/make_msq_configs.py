"""Convert the reference's shipped MSQ results (msq_results/figure1{c,d}/*.pt: qdict + merge_info, BASELINE.json
configs[2]) to JSON fixtures under configs/ so that bench.py / tests can use them on the GPU box, where /root/reference
does not exist.  Run in the build container:  python tools/make_msq_configs.py"""
import glob, json, os
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/msq_results"
for fig in ("figure1c", "figure1d"):
    (qpath,) = [p for p in glob.glob(os.path.join(REF, fig, "*.pt")) if not p.endswith("_merge_info.pt")]
    qdict = torch.load(qpath, weights_only=False)
    merge = torch.load(qpath[:-3] + "_merge_info.pt", weights_only=False)
    out = {"source": os.path.relpath(qpath, "/root/reference"),
           "qdict": {k: (list(v) if not isinstance(v, str) else [v, "0"]) for k, v in qdict.items()},
           "merge_info": [list(m) for m in merge]}
    with open(os.path.join(ROOT, "configs", fig + ".json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(fig, len(out["qdict"]), "entries")
