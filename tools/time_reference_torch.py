"""Build container only (needs /root/reference): the reference's OWN torch CPU decode for an even-KV layer --
lib/utils/kernel_decompress.py:decode_compressed with the expanded LUT of lib/codebook/bitshift.py:quantlut_sym, then
x.float() @ W.float().T -- timed beside the C port (oracle/qp_cref.c) that bench.py's CPU arm uses, on the same host, same
shape, same thread count, after checking that both decode the same weights.  BASELINE.md section 3 promised this number.

    python tools/time_reference_torch.py > profiles/r02_reference_torch_cpu.log
"""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import REF, _import_reference  # noqa: E402
import bench  # noqa: E402  (for the C port loader)

_import_reference()
from lib.codebook.bitshift import quantlut_sym  # noqa: E402
from lib.utils.kernel_decompress import decode_compressed  # noqa: E402

torch.set_num_threads(os.cpu_count())
M, K, KV, S = 4096, 4096, 6, 9
rng = np.random.default_rng(0)
tlut = torch.load(f"{REF}/assets/lut_cache/kmeans_{S}_2.pt").half()
buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
x = rng.standard_normal((1, K)).astype(np.float16)
exp = quantlut_sym(tlut, 16, S)
fn = getattr(decode_compressed, "_torchdynamo_orig_callable", decode_compressed)
best_dec, best_mm = float("inf"), float("inf")
for _ in range(3):
    t0 = time.perf_counter()
    W = fn(16, S, KV // 2, 1, M, K, torch.from_numpy(buf.view(np.uint16).copy()), exp)
    t1 = time.perf_counter()
    y = torch.from_numpy(x).float() @ W.float().T
    t2 = time.perf_counter()
    best_dec, best_mm = min(best_dec, t1 - t0), min(best_mm, t2 - t1)
lib = bench._cref()
vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
tl = tlut.numpy()
out = np.zeros((1, M), np.float32)
best_c = float("inf")
for _ in range(3):
    out[:] = 0
    t0 = time.perf_counter()
    lib.qp_cref_tcq(vp(buf), vp(tl), M, K, KV, S, vp(x), 1, K, 0, 0, M, vp(out), None, 0)
    best_c = min(best_c, time.perf_counter() - t0)
err = float(np.linalg.norm(out - y.numpy()) / np.linalg.norm(y.numpy()))
print(f"host: {os.cpu_count()} cores, torch {torch.__version__} (eager), layer {M}x{K} tcq_{KV} (S = {S}), bs = 1, best of 3")
print(f"reference torch path : decode_compressed {best_dec * 1e3:8.1f} ms + matmul {best_mm * 1e3:6.1f} ms = {(best_dec + best_mm) * 1e3:8.1f} ms")
print(f"C port (bench.py arm): decode + matvec   {best_c * 1e3:8.1f} ms on {int(lib.qp_cref_threads())} threads   -> {(best_dec + best_mm) / best_c:.1f}x faster than the torch path")
print(f"same result: rel-L2 of the two outputs {err:.2e}")
