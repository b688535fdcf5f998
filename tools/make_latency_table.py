"""B200 latency-coefficient table in the schema the reference's MSQ solver reads (solve_lat_const.py:113-123,219-221:
`assets/{model_key}_latency_coeffs_{nodename}.pt`): a dict

    "{q|k|v|o|u|g|d|qk|kv|qv|qkv|ug}_{quantizer_str}_{True|False}" -> seconds (float)     588 entries
    "constant"                                                      -> seconds (0-dim tensor)

where the boolean is the SIMT-layout flag (ldlq quantizers only) and the solver models a token's latency as
constant + sum over layers of the chosen projections' coefficients.  Here a coefficient is the device time of that
(possibly merged) projection's bs=1 GEMV at the Llama-3.1-8B shapes, graph-replayed over enough distinct weight buffers to
defeat the L2 (as bench.py does); `constant` is what is left of a measured decode step of the uniform tcomb_6_7 model after
subtracting its 32 x (qkv + o + ug + d) coefficients (glue kernels, attention, lm_head, sampling).

    python tools/make_latency_table.py [--out profiles/3_8b_latency_coeffs_b200] [--quick]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette.decode import LLAMA31_8B, DecodeRunner, _Proj, uniform_qdict  # noqa: E402

H, KVD, I = 4096, 1024, 14336
SHAPES = {"q": (H, H), "k": (KVD, H), "v": (KVD, H), "o": (H, H), "u": (I, H), "g": (I, H), "d": (H, I),
          "qk": (H + KVD, H), "kv": (2 * KVD, H), "qv": (H + KVD, H), "qkv": (H + 2 * KVD, H), "ug": (2 * I, H)}
QUANTIZERS = ([f"tcq_{k}_none_0.9" for k in range(3, 11)] + [f"tcomb_{k}_{k + 1}_0.5_none_0.9" for k in range(3, 10)] +
              [f"ldlq_1_{b}_none_1.0" for b in range(2, 9)] + [f"ldlq_2_{b}_none_1.0" for b in range(3, 13)])


def time_rotation(projs, x, out, iters):
    st = torch.cuda.current_stream()
    for p in projs:
        p.launch(out.data_ptr(), x.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for p in projs:
            p.launch(out.data_ptr(), x.data_ptr(), torch.cuda.current_stream().cuda_stream)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / (iters * len(projs))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "3_8b_latency_coeffs_b200"))
    ap.add_argument("--quick", action="store_true", help="2 quantizers only (smoke run)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    table, t0 = {}, time.time()
    quantizers = QUANTIZERS if not a.quick else ["tcomb_6_7_0.5_none_0.9", "ldlq_2_8_none_1.0"]
    for qs in quantizers:
        for simt in ([False, True] if qs.startswith("ldlq") else [False]):
            for key, (M, K) in SHAPES.items():
                p0 = _Proj(qs, "1" if simt else "0", K, M, dev, gen)
                ncopy = max(3, min(64, int(280e6 // max(p0.weight_bytes, 1)) + 1))
                projs = [p0] + [_Proj(qs, "1" if simt else "0", K, M, dev, gen) for _ in range(ncopy - 1)]
                x = torch.randn(K, device=dev).half()
                out = torch.zeros(M, dtype=torch.float32, device=dev)
                t = time_rotation(projs, x, out, iters=max(2, 300 // ncopy))
                table[f"{key}_{qs}_{simt}"] = float(t)
                del projs, p0
            torch.cuda.empty_cache()
        print(f"{qs}: d {table[f'd_{qs}_False'] * 1e6:.2f} us  ug {table[f'ug_{qs}_False'] * 1e6:.2f} us  ({time.time() - t0:.0f} s)", flush=True)
    # constant: a measured decode step of the uniform model minus its projections
    qs = "tcomb_6_7_0.5_none_0.9"
    r = DecodeRunner(LLAMA31_8B, uniform_qdict(LLAMA31_8B, qs), [["merge_qkv", "merge_ug"]] * 32, max_seq=128, seed=0)
    r.capture()
    r.reset(1)
    for _ in range(8):
        r.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        r.step()
    e1.record()
    torch.cuda.synchronize()
    step_s = e0.elapsed_time(e1) * 1e-3 / 40
    layer_s = sum(table[f"{k}_{qs}_False"] for k in ("qkv", "o", "ug", "d"))
    table["constant"] = torch.tensor(step_s - 32 * layer_s)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    torch.save(table, a.out + ".pt")
    meta = {"device": torch.cuda.get_device_name(0), "entries": len(table), "decode_step_s": step_s, "layer_gemv_s": layer_s,
            "constant_s": float(table["constant"]), "seconds_to_build": time.time() - t0,
            "how": "bs=1 GEMV per (projection key, quantizer, simt) at the Llama-3.1-8B shapes, CUDA-graph replay over >= 280 MB of "
                   "distinct weight buffers, CUDA events; constant = measured uniform tcomb_6_7 decode step - 32 * (qkv + o + ug + d)"}
    json.dump({"meta": meta, "table_us": {k: (float(v) * 1e6) for k, v in table.items()}}, open(a.out + ".json", "w"), indent=0)
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
