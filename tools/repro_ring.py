import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200")); sys.path.insert(0, ROOT)
from qpalette import ops
from qpalette._cabi import SPLIT_IN
from oracle import qp_oracle as O
M, K = int(sys.argv[1]), int(sys.argv[2])
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(0)
b1 = rng.integers(0, 256, size=M * (K // 2) * 6 // 16, dtype=np.uint8)
b2 = rng.integers(0, 256, size=M * (K // 2) * 7 // 16, dtype=np.uint8)
tl = (rng.standard_normal((512, 2)) * 0.9).astype(np.float16)
x = rng.standard_normal((bs, K)).astype(np.float16)
d = lambda a: torch.from_numpy(a).cuda()
out = ops.tcq_gemv(d(b1), d(x), d(tl), M, K, 9, 6, d(b2), 7, SPLIT_IN, K // 2)
torch.cuda.synchronize()
ref = O.gemv_ref(O.tcq_decode_combt(b1, b2, tl, M, K, 6, 7, 9), x)
o = out.cpu().numpy()
print(M, K, bs, "rel-L2", np.linalg.norm(o - ref) / np.linalg.norm(ref))
