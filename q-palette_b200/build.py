"""Build libqpalette.so (sm_100a) in-tree with plain nvcc.  Used by __graft_entry__.build() and by hand:

    python q-palette_b200/build.py [--force]

The library is a C-ABI shared object (include/qpalette.h); it has no torch / pybind dependency, so it builds in
about a minute and travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "qpalette", "libqpalette.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# no --use_fast_math: measured round 2 (profiles/r02_fastmath_delta.log) it changes neither the decode rate (498.2 vs 498.9 tok/s)
# nor the distance to the float64 restatement (5.1e-4 vs 5.5e-4, below the run-to-run noise of the fp32 atomics), so the
# library is built with IEEE division / sqrt and denormals; the kernels call __expf where a fast exponential is meant
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v"]
# tuning knobs of the streaming GEMV kernels (see csrc/gemv_common.cuh); override for experiments
for _k in ("QP_GEMV_THREADS", "QP_PROFILE_PHASES", "QP_HASH_ADD", "QP_TC_TCQ_STRIDE", "QP_GEMV_CTAS", "QP_TCQ_FOLD", "QP_TC_BACKOFF", "QP_TCQ_STRIDE9", "QP_GEMV2_DEPTH", "QP_FAST_BUILD", "QP_MMA_DEPTH4", "QP_MMA_THREADS4", "QP_REFILL_LATE", "QP_TRIP_CHECK", "QP_XP_FINAL32"):
    if os.environ.get(_k):
        FLAGS.append(f"-D{_k}={os.environ[_k]}")
if os.environ.get("QP_FAST_MATH"):  # experiment: what --use_fast_math does to SiLU / RMSNorm / softmax
    FLAGS.append("--use_fast_math")
if os.environ.get("QP_LIB_SUFFIX"):
    OUT = OUT.replace("libqpalette.so", f"libqpalette{os.environ['QP_LIB_SUFFIX']}.so")
    OBJ = OBJ + os.environ["QP_LIB_SUFFIX"]
SOURCES = ["qp_api.cu", "tcq_kernels.cu", "lut_kernels.cu", "simt_kernels.cu", "had_kernels.cu", "decode_kernels.cu",
           "gemm_tc_kernels.cu"]


def _digest(path):
    h = hashlib.sha1()
    for name in sorted(os.listdir(CSRC)) + ["../../include/qpalette.h"]:
        p = os.path.join(CSRC, name)
        if os.path.isfile(p) and (name.endswith((".cuh", ".h")) or os.path.abspath(p) == os.path.abspath(path)):
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    stamp = obj + ".sha1"
    d = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == d:
        return obj, "", False
    r = subprocess.run([NVCC, *FLAGS, "-c", path, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(d)
    return obj, r.stderr, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [r[0] for r in results]
    if verbose:
        for _, log, _ in results:
            sys.stderr.write(log)
    if any(r[2] for r in results) or not os.path.exists(OUT):
        # cudart is linked statically (nvcc default): no dependency on a libcudart.so being on the loader path
        r = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
        for _, log, fresh in results:
            if fresh:
                f.write(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
