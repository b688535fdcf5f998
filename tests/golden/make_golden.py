"""Generate tests/golden/*.npz by running the REFERENCE's own python on CPU (build container only).

    python tests/golden/make_golden.py            # needs /root/reference; writes next to this file

The fixtures pin oracle/qp_oracle.py (and, through it, the CUDA path).  Every array below is the output of an
unmodified reference function; nothing from the oracle is used to produce them.  Reference modules are imported
with `glog` and `fast_hadamard_transform` stubbed (neither is installed; the latter is third-party, see the oracle
header) and `Tensor.cuda()` neutralised (packers hard-code `.cuda()`: quant_op.py:76,85,158).
"""
import os
import sys
import types

os.environ["TORCHDYNAMO_DISABLE"] = "1"  # reference decoders are @torch.compile; run them eagerly

import numpy as np
import scipy.linalg
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _import_reference():
    g = types.ModuleType("glog")
    g.info = print
    sys.modules["glog"] = g
    f = types.ModuleType("fast_hadamard_transform")

    def hadamard_transform(x, scale=1.0):
        n = x.shape[-1]
        H = torch.tensor(scipy.linalg.hadamard(n), dtype=torch.float32)
        return (x.float() @ H.T * scale).to(x.dtype)

    f.hadamard_transform = hadamard_transform
    sys.modules["fast_hadamard_transform"] = f
    sys.path.insert(0, REF)
    os.chdir(REF)
    torch.Tensor.cuda = lambda self, *a, **k: self


def tail_biting_states(rng, n_tiles, KV):
    nb = 128 * KV
    bits = rng.integers(0, 2, size=(n_tiles, nb), dtype=np.uint8)
    pos = (np.arange(128)[:, None] * KV + np.arange(16)[None, :]) % nb
    w = bits[:, pos].astype(np.uint32)
    return (w << np.arange(15, -1, -1, dtype=np.uint32)).sum(-1).astype(np.int32)


def main():
    _import_reference()
    from lib.codebook.bitshift import bitshift_codebook, quantlut_sym
    from lib.utils.kernel_decompress import decode_compressed
    from lib.quantizer import quant_op
    from lib.utils import matmul_had
    from lib.utils.mem_op import get_quant_info

    rng = np.random.default_rng(1234)
    torch.manual_seed(1234)
    out = {}

    # ---- shipped TCQ codebooks (assets/lut_cache/kmeans_{S}_2.pt), as fp16, and quantlut_sym of them
    for S in (9, 10, 11):
        tlut = torch.load(f"{REF}/assets/lut_cache/kmeans_{S}_2.pt").half()
        out[f"tlut_{S}"] = tlut.numpy()
        exp = quantlut_sym(tlut, 16, S)
        out[f"quantlut_sym_{S}_sample_idx"] = np.arange(0, 65536, 97, dtype=np.int64)
        out[f"quantlut_sym_{S}_sample"] = exp[::97].numpy()
        out[f"quantlut_sym_{S}_sum"] = np.array([exp.double().sum().item(), exp.double().abs().sum().item()])

    # ---- TCQ pack_trellis (bitshift.py:296-329) for every KV, and decode_compressed (kernel_decompress.py) for even KV
    M, K = 64, 128
    for KV in range(2, 11):
        S = 9 if KV <= 8 else KV + 1
        tlut = torch.from_numpy(out[f"tlut_{S}"])
        cb = bitshift_codebook(L=16, KV=KV, V=2, tlut_bits=S, decode_mode="quantlut_sym", tlut=tlut)
        states = tail_biting_states(rng, (M // 16) * (K // 16), KV)
        packed = cb.pack_trellis(torch.from_numpy(states))  # uint16 (tiles, 8*KV), BEFORE the kernel swizzle
        out[f"tcq_states_{KV}"] = states.astype(np.uint16)
        out[f"tcq_pack_trellis_{KV}"] = packed.view(torch.int16).numpy()
        out[f"tcq_recons_{KV}"] = cb.recons(torch.from_numpy(states)).numpy()  # (2, tiles, 128) fp16
        if KV % 2 == 0:
            # random bytes are valid input: decode with the reference's torch decoder
            buf = torch.from_numpy(rng.integers(0, 65536, size=M * K * KV // 32, dtype=np.uint16).view(np.int16))
            exp = quantlut_sym(tlut, 16, S)
            W = decode_compressed(16, S, KV // 2, 1, M, K, buf.view(torch.uint16), exp)
            out[f"tcq_buf_{KV}"] = buf.numpy()
            out[f"tcq_decode_compressed_{KV}"] = W.numpy()

    # ---- VQ / SQ tensor-core layout: pack_qweight + dequantize_mat_sq_inds[_vec2] (quant_op.py)
    M, K = 64, 128
    for vec, Rs in ((1, range(2, 9)), (2, range(2, 13))):
        for R in Rs:
            Q = torch.from_numpy(rng.integers(0, 1 << R, size=(M, K // vec), dtype=np.int64))
            packed = quant_op.pack_qweight(Q, vec, R)
            fn = quant_op.dequantize_mat_sq_inds if vec == 1 else quant_op.dequantize_mat_sq_inds_vec2
            back = fn(packed, M, K, R)
            assert (back == Q).all()
            out[f"lut_tc_codes_{vec}_{R}"] = Q.numpy().astype(np.int32)
            out[f"lut_tc_packed_{vec}_{R}"] = packed.view(torch.int32).numpy()

    # ---- SIMT layouts: pack_qweight_sq_simt / pack_qweight_vq_simt (+ a ragged last chunk), convert_tensor_core_to_simt
    for vec, Rs, Ks in ((1, (2, 3, 4, 5, 8), (1024, 1024 + 256)), (2, (3, 6, 8, 12), (2048, 2048 + 512))):
        for R in Rs:
            for K in Ks:
                M = 8
                Q = torch.from_numpy(rng.integers(0, 1 << R, size=(M, K // vec), dtype=np.int64))
                if vec == 1:
                    packed = quant_op.pack_qweight_sq_simt(Q, R)
                else:
                    packed = quant_op.pack_qweight_vq_simt(Q, R, vec, R)
                out[f"simt_codes_{vec}_{R}_{K}"] = Q.numpy().astype(np.int32)
                out[f"simt_packed_{vec}_{R}_{K}"] = packed.numpy().view(np.int32).reshape(M, -1)
    for vec, R in ((1, 4), (2, 6)):
        M, K = 32, 2048
        Q = torch.from_numpy(rng.integers(0, 1 << R, size=(M, K // vec), dtype=np.int64))
        tc = quant_op.pack_qweight(Q, vec, R)
        simt = quant_op.convert_tensor_core_to_simt(tc, M, K, vec, R, code_n=R)
        out[f"conv_tc_{vec}_{R}"] = tc.view(torch.int32).numpy()
        out[f"conv_simt_{vec}_{R}"] = simt.numpy().view(np.int32).reshape(M, -1)

    # ---- Hadamard: had28 literal, pure-torch matmul_hadU / matmul_hadUt, and the *_cuda call forms
    out["had28"] = matmul_had.get_had28().numpy()
    for n in (1024, 4096, 14336):
        x = torch.randn(2, n, dtype=torch.float32)
        out[f"had_x_{n}"] = x.numpy()
        out[f"had_U_{n}"] = matmul_had.matmul_hadU(x).numpy()
        out[f"had_Ut_{n}"] = matmul_had.matmul_hadUt(x).numpy()
        hadK, Kf = matmul_had.get_hadK(n)
        hadK_T = hadK.T.contiguous() if hadK is not None else None
        out[f"had_cuda_T_{n}"] = matmul_had.matmul_hadU_cuda(x, hadK_T, Kf).numpy()
        out[f"had_head_cuda_T_{n}"] = matmul_had.matmul_hadU_head_cuda(x, hadK_T, Kf, n).numpy()

    # ---- quantizer-string grammar (mem_op.py:271-307)
    for qs in ("tcq_6_none_0.9", "tcq_9_none_0.9", "tcq_10_none_0.9", "tcomb_6_7_0.5_none_0.9",
               "tcomb_9_10_0.5_none_0.9", "ldlq_2_12_none_1.0", "ldlq_1_6_none_1.0"):
        qi = get_quant_info(qs)
        flat = [qi.get("tlut_bits", -1), qi.get("lut_bits", -1), qi.get("vec_sz", -1)]
        kv = qi.get("KV", -1)
        flat += list(kv) if isinstance(kv, list) else [kv, -1]
        out[f"qinfo_{qs}"] = np.array(flat, dtype=np.int64)

    # ---- SIMT layout with 4-wide vectors (ours_lib::vq_pack_*_simt_*_4_*, lib/linear/__init__.py:383-420); its own generator
    # so that the arrays above keep their values
    rng4 = np.random.default_rng(404)
    for R in (6, 8, 10, 12):
        for K in (4096, 4096 + 1024):
            M = 8
            Q = torch.from_numpy(rng4.integers(0, 1 << R, size=(M, K // 4), dtype=np.int64))
            packed = quant_op.pack_qweight_vq_simt(Q, R, 4, R)
            out[f"simt_codes_4_{R}_{K}"] = Q.numpy().astype(np.int32)
            out[f"simt_packed_4_{R}_{K}"] = packed.numpy().view(np.int32).reshape(M, -1)

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
