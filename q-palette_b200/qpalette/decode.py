"""bs=1 decode loop for Llama-shaped models built from Q-Palette quantized layers (the path eval/measure_latency.py times).

The reference wraps a HF Llama with `torch.compile(mode="max-autotune", fullgraph=True)` + CUDA graphs
(eval/measure_latency.py:188-273).  Here the decode step is an explicit, flat list of libqpalette kernel launches
(4 quantized GEMVs + 5 fused glue kernels per layer, see csrc/decode_kernels.cu) captured ONCE into a CUDA graph;
token and position live in device memory, so the same graph is replayed for every token.

Layers come from a `qdict` ("{layer}_{self_attn.q_proj|...|mlp.down_proj}" -> quantizer_str or (quantizer_str, simt_flag))
and a `merge_info` (per-layer list from {merge_qkv, merge_qk, merge_kv, merge_qv, merge_ug}), exactly the reference's
formats (eval/measure_latency_merge_simt.py:60-71, solve_lat_const.py:152-162), with the reference's `--dummy` random
initialisation (lib/utils/mem_op.py:198-269) since no checkpoints are reachable offline.

Row sharding (tensor parallel over output rows): with world_size > 1 every rank keeps the rows
[rank, rank+1) * M/world of each projection (a zero-copy slice of the strip-major packed buffers) and the partial outputs
are all-gathered with NCCL at the four layer boundaries.
"""
import ctypes
import math
from dataclasses import dataclass, field

import torch

from . import _cabi
from ._cabi import FLAG_ACCUMULATE, SPLIT_IN, SPLIT_NONE, check, lib
from .linear.incoherent_linear import rope_inv_freq
from .shard import shard_plan, shard_rows
from .utils.mem_op import get_dummy_quant_results, get_quant_info


@dataclass
class LlamaShape:
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    vocab_size: int = 128256
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: dict = field(default_factory=lambda: dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0,
                                                            high_freq_factor=4.0, original_max_position_embeddings=8192))
    hidden_act: str = "silu"
    attention_dropout: float = 0.0
    max_position_embeddings: int = 131072
    _name_or_path: str = "meta-llama/Llama-3.1-8B"

    @property
    def head_dim(self):
        return self.hidden_size // self.num_attention_heads


LLAMA31_8B = LlamaShape()
LLAMA31_70B = LlamaShape(hidden_size=8192, intermediate_size=28672, num_hidden_layers=80, num_attention_heads=64,
                         num_key_value_heads=8, _name_or_path="meta-llama/Llama-3.1-70B")


def uniform_qdict(shape, quantizer_str, simt="0"):
    keys = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.up_proj",
            "mlp.gate_proj", "mlp.down_proj"]
    return {f"{i}_{k}": (quantizer_str, simt) for i in range(shape.num_hidden_layers) for k in keys}


class _Proj:
    """one quantized projection (possibly a merged one): packed buffers + what the C ABI needs to run its GEMV."""

    def __init__(self, quantizer_str, simt, in_features, out_features, device, gen, rank=0, world=1, group_sizes=None):
        self.qs, self.K = quantizer_str, in_features
        qi = get_quant_info(quantizer_str)
        self.kind = qi["quantizer"]
        info = get_dummy_quant_results(None, None, quantizer_str, device=device, generator=gen, in_features=in_features,
                                       out_features=out_features)["linear_info"]
        self.M_full = out_features
        # row shard: keep rows [rank, rank+1) * M/world of every merged member (zero-copy slices: strip-major layout)
        rows = shard_plan(group_sizes or [out_features], rank, world)
        self.M = sum(n for _, n in rows)
        shard = lambda t: shard_rows(t, out_features, rows)

        self.bufs = []
        if self.kind == "tcq_ldlq":
            self.S, self.KV1, self.KV2, self.split, self.part1 = qi["tlut_bits"], qi["KV"], 0, SPLIT_NONE, 0
            self.codes1, self.codes2 = shard(info["trellis"]), None
            self.lut = info["tlut"]
            self.bits_per_weight = qi["KV"] / 2
        elif self.kind == "combt_ldlq":
            self.S, (self.KV1, self.KV2), self.split, self.part1 = qi["tlut_bits"], qi["KV"], SPLIT_IN, in_features // 2
            self.codes1, self.codes2 = shard(info["trellis1"]), shard(info["trellis2"])
            self.lut = info["tlut"]
            self.bits_per_weight = (qi["KV"][0] + qi["KV"][1]) / 4
        elif self.kind in ("vq_ldlq", "vq"):
            self.bits, self.vec = qi["lut_bits"], qi["vec_sz"]
            self.lut = info["lut"]
            self.simt = bool(int(simt)) and self.vec <= 2
            qw = shard(info["qweight"])
            if self.simt:
                out = torch.empty_like(qw)
                check(lib().qp_convert_tc_to_simt(out.data_ptr(), qw.data_ptr(), self.M, in_features, self.bits, self.vec,
                                                  torch.cuda.current_stream().cuda_stream))
                qw = out
            self.codes1, self.codes2 = qw, None
            self.bits_per_weight = self.bits / self.vec
        else:
            raise ValueError(f"unsupported quantizer {quantizer_str}")
        self.weight_bytes = self.codes1.numel() * self.codes1.element_size() + \
            (self.codes2.numel() * self.codes2.element_size() if self.codes2 is not None else 0)

    def launch(self, out_ptr, x_ptr, stream):
        L = lib()
        if self.kind in ("tcq_ldlq", "combt_ldlq"):
            check(L.qp_tcq_gemv(out_ptr, self.codes1.data_ptr(), self.codes2.data_ptr() if self.codes2 is not None else None,
                                x_ptr, self.lut.data_ptr(), self.M, self.K, 1, self.S, self.KV1, self.KV2, self.split,
                                self.part1, FLAG_ACCUMULATE, stream))
        elif self.simt:
            check(L.qp_simt_gemv(out_ptr, self.codes1.data_ptr(), x_ptr, self.lut.data_ptr(), self.M, self.K, 1, self.bits,
                                 self.vec, 1, stream))
        else:
            check(L.qp_lut_gemv(out_ptr, self.codes1.data_ptr(), x_ptr, self.lut.data_ptr(), self.M, self.K, 1, self.bits,
                                self.vec, FLAG_ACCUMULATE, stream))


    def can_fuse(self):
        return self.kind in ("tcq_ldlq", "combt_ldlq") or not self.simt

    def launch_fused(self, out_ptr, xprod, stream):
        """GEMV with the activation glue computed in its prologue (xprod: _cabi.XProd)"""
        import ctypes
        L = lib()
        if self.kind in ("tcq_ldlq", "combt_ldlq"):
            check(L.qp_tcq_gemv_fused(out_ptr, self.codes1.data_ptr(),
                                      self.codes2.data_ptr() if self.codes2 is not None else None, ctypes.addressof(xprod),
                                      self.lut.data_ptr(), self.M, self.K, self.S, self.KV1, self.KV2, self.split,
                                      self.part1, stream))
        else:
            check(L.qp_lut_gemv_fused(out_ptr, self.codes1.data_ptr(), ctypes.addressof(xprod), self.lut.data_ptr(), self.M,
                                      self.K, self.bits, self.vec, stream))


def silu_grid_supported(I):
    """shapes the multi-CTA SiLU*mul + Hadamard kernel is instantiated for: I = Kf * R * 512 (csrc/decode_kernels.cu)"""
    for kf, rs in ((28, (1, 2)), (1, (8, 16, 32))):
        if I % (kf * 512) == 0 and I // (kf * 512) in rs:
            return True
    return False


class DecodeRunner:
    def __init__(self, shape=LLAMA31_8B, qdict=None, merge_info=None, max_seq=512, device="cuda", seed=0, rank=0,
                 world=1, process_group=None, num_layers=None, random_scales=True, fused=True, p2p=True):
        """world > 1: rows of every projection are sharded over the ranks of `process_group`.  p2p=True gathers
        at the four layer boundaries by NVLink peer stores fused into the consumer kernel (qp_fused_norm_had_xchg);
        p2p=False uses ncclAllGather (the baseline)."""
        self.shape, self.dev, self.rank, self.world, self.pg = shape, torch.device(device), rank, world, process_group
        self.p2p = p2p and world > 1
        self.L = num_layers or shape.num_hidden_layers
        self.max_seq = max_seq
        self.fused = fused   # GEMV-prologue fusion of the glue (row-sharded: only with the peer exchange, decided below)
        qdict = qdict or uniform_qdict(shape, "tcomb_6_7_0.5_none_0.9")
        merge_info = merge_info or [["merge_qkv", "merge_ug"]] * shape.num_hidden_layers
        g = torch.Generator(device=self.dev)
        g.manual_seed(seed)
        H, kvd, I, V = shape.hidden_size, shape.num_key_value_heads * shape.head_dim, shape.intermediate_size, shape.vocab_size
        self.H, self.kvd, self.I = H, kvd, I
        assert shape.num_attention_heads % world == 0 and shape.num_key_value_heads % world == 0 or world == 1
        f16 = dict(dtype=torch.float16, device=self.dev)

        def rnd(*s, scale=1.0):
            return (torch.randn(*s, generator=g, dtype=torch.float32, device=self.dev) * scale).to(torch.float16)

        def signs(n):
            return ((torch.randn(n, generator=g, device=self.dev) > 0).float() * 2 - 1).to(torch.float16)

        def wscale(n):
            # per-row scale so that quantized weights (unit variance codebooks) act like 1/sqrt(K)-scaled matrices
            if not random_scales:
                return torch.ones(n, **f16)
            return (torch.rand(n, generator=g, device=self.dev) * 0.5 + 0.75).to(torch.float16)

        self.embed = rnd(V, H, scale=1.0)
        self.lm_head = rnd(V, H, scale=H ** -0.5)  # fp16, NOT quantized (eval/measure_latency.py keeps lm_head in fp16)
        self.final_norm = torch.ones(H, **f16)
        self.inv_freq = rope_inv_freq(shape, self.dev)
        self.layers = []
        self.weight_bytes = 0

        def entry(i, key):
            v = qdict[f"{i}_{key}"]
            return (v, "0") if isinstance(v, str) else (v[0], v[1])

        for i in range(self.L):
            merges = merge_info[i] if merge_info is not None else []
            ly = {}
            q, k, v, o = (entry(i, f"self_attn.{n}_proj") for n in "qkvo")
            up, gate, down = (entry(i, f"mlp.{n}_proj") for n in ("up", "gate", "down"))
            mk = lambda e, kin, m, sizes=None: _Proj(e[0], e[1], kin, m, self.dev, g, rank, world, sizes)
            # attention projections: list of (proj, offset into the [q|k|v] accumulator of THIS rank)
            Hq, Hk = H // world, kvd // world
            if "merge_qkv" in merges:
                ly["qkv"] = [(mk(q, H, H + 2 * kvd, [H, kvd, kvd]), 0)]
            elif "merge_qk" in merges:
                ly["qkv"] = [(mk(q, H, H + kvd, [H, kvd]), 0), (mk(v, H, kvd), Hq + Hk)]
            elif "merge_kv" in merges:
                ly["qkv"] = [(mk(q, H, H), 0), (mk(k, H, 2 * kvd, [kvd, kvd]), Hq)]
            elif "merge_qv" in merges:
                # accumulator (and Wscale) order q | v | k for this layer, as the reference stores Wscale_qkv for merge_qv
                # (lib/linear/incoherent_linear.py:211-213); the attention kernel is told through `qvk_order`
                ly["qkv"] = [(mk(q, H, H + kvd, [H, kvd]), 0), (mk(k, H, kvd), Hq + Hk)]
                ly["qvk"] = 1
            else:
                ly["qkv"] = [(mk(q, H, H), 0), (mk(k, H, kvd), Hq), (mk(v, H, kvd), Hq + Hk)]
            ly["o"] = mk(o, H, H)
            Il = I // world
            if "merge_ug" in merges:
                ly["ug"] = [(mk(up, H, 2 * I, [I, I]), 0)]
            else:
                ly["ug"] = [(mk(up, H, I), 0), (mk(gate, H, I), Il)]
            ly["down"] = mk(down, I, H)
            ly["norm1"], ly["norm2"] = torch.ones(H, **f16), torch.ones(H, **f16)
            ly["SU_qkv"], ly["SU_o"], ly["SU_ug"], ly["SU_dp"] = signs(H), signs(H), signs(H), signs(I)
            # Wscale ~ 1/sqrt(K)/s keeps activations O(1) through random-init layers (Wscale is per output row).
            # Drawn at FULL width on every rank (same generator state everywhere) and sliced like the weight rows, so a
            # row-sharded run reproduces the world = 1 model exactly.
            def rows_of(full, sizes):
                return torch.cat([full[r0:r0 + n] for r0, n in shard_plan(sizes, rank, world)]).contiguous()

            ly["W_qkv"] = rows_of((wscale(H + 2 * kvd) * (H ** -0.5)).to(torch.float16), [H, kvd, kvd])
            ly["W_o"] = rows_of((wscale(H) * (H ** -0.5)).to(torch.float16), [H])
            ly["W_ug"] = rows_of((wscale(2 * I) * (H ** -0.5)).to(torch.float16), [I, I])
            ly["W_dp"] = rows_of((wscale(H) * (I ** -0.5)).to(torch.float16), [H])
            ly["kc"] = torch.zeros((max_seq, shape.num_key_value_heads // world, shape.head_dim), **f16)
            ly["vc"] = torch.zeros_like(ly["kc"])
            for p in [pp for pp, _ in ly["qkv"]] + [ly["o"]] + [pp for pp, _ in ly["ug"]] + [ly["down"]]:
                self.weight_bytes += p.weight_bytes
            self.layers.append(ly)

        z32 = dict(dtype=torch.float32, device=self.dev)
        Hq, Hk, Il = H // world, kvd // world, I // world
        self.h = torch.zeros(H, **f16)
        self.h2 = torch.zeros(H, **f16)          # ping-pong partner of h for the fused prologues
        self.x_h = torch.zeros(H, **f16)
        self.x_i = torch.zeros(I, **f16)
        # split-KV attention: partial softmaxes + one ticket per head, shared by all layers (zeroed once; launches leave it zero)
        nb = int(lib().qp_rope_attention_scratch_bytes(shape.num_attention_heads // world, shape.head_dim, max_seq))
        self.attn_scratch = torch.zeros(nb, dtype=torch.uint8, device=self.dev) if nb else None
        self.region = None
        if self.p2p:
            # the four gathered buffers live in a region every peer maps; site = 4 * layer + {attn, o, act, down}
            from .peer import PeerRegion
            try:
                # ll_*: receive buffers of the low-latency gather in front of the fused GEMV prologues (4 bytes per element)
                self.region = PeerRegion([("attn", H * 2), ("acc_o", H * 4), ("act", I * 2), ("acc_dn", H * 4), ("z", I * 4),
                                          ("ll_attn", H * 4), ("ll_acc_o", H * 4), ("ll_acc_dn", H * 4)],
                                         4 * self.L + 4, rank, world, process_group, self.dev)
            except Exception as ex:  # e.g. CUDA IPC not permitted between these processes
                self.region = None
                print(f"[qpalette] rank {rank}: peer exchange unavailable ({ex}); using ncclAllGather", flush=True)
            # every rank must take the same path
            ok = torch.tensor([1 if self.region is not None else 0], device=self.dev)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=process_group)
            if int(ok.item()) == 0:
                self.region, self.p2p = None, False
        if self.p2p:
            self.attn = self.region.tensor("attn", torch.float16)
            self.acc_o = self.region.tensor("acc_o", torch.float32)
            self.act = self.region.tensor("act", torch.float16)
            self.acc_dn = self.region.tensor("acc_dn", torch.float32)
        else:
            self.attn = torch.zeros(H, **f16)       # full attention output (all-gathered when sharded)
            self.acc_o = torch.zeros(H, **z32)       # full width; each rank fills its slice, then all-gather
            self.acc_dn = torch.zeros(H, **z32)
            self.act = torch.zeros(I, **f16)
        self.attn_loc = torch.zeros(Hq, **f16)
        self.acc_qkv = torch.zeros(Hq + 2 * Hk, **z32)
        self.acc_ug = torch.zeros(2 * Il, **z32)
        self.act_loc = torch.zeros(Il, **f16)    # sharded silu(gate)*up
        self.xf = torch.zeros(H, **f16)
        self.logits = torch.zeros(V, **z32)
        self.token = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.history = torch.zeros(max_seq, dtype=torch.int32, device=self.dev)
        self.scratch = torch.zeros(4096, dtype=torch.uint8, device=self.dev)
        self.sync = torch.zeros(4, dtype=torch.int32, device=self.dev)  # ticket of the multi-CTA SiLU*mul/Hadamard kernel
        self.silu_grid = silu_grid_supported(self.I)
        # row-sharded form: the I / 512 blocks must split evenly over the ranks (and the shape be instantiated)
        self.silu_grid_tp = (self.p2p and self.I % (512 * world) == 0 and
                             any(self.I == kf * r * 512 for kf, r in ((28, 1), (28, 2), (1, 8), (1, 16))))
        if world > 1 and not (self.p2p and self.silu_grid_tp):
            self.fused = False  # the NCCL baseline and shapes without the multi-CTA exchange kernel keep the un-fused list
        self.graph = None
        self.steps_done = 0          # host-side count of positions written since reset(): the KV cache holds max_seq rows
        self.lm_head_bytes = self.lm_head.numel() * 2
        self.launches_per_step = 0
        self._prepare_full_scales()
        for ly in self.layers:  # the fused prologue needs one tensor-core-layout member per group
            for grp in (ly["qkv"], [(ly["o"], 0)], ly["ug"]):
                if not any(pr.can_fuse() for pr, _ in grp):
                    self.fused = False

    # -----------------------------------------------------------------------------------------------------------------
    def _step(self):
        if self.fused:
            return self._step_fused() if self.world == 1 else self._step_fused_tp()
        return self._step_unfused()

    def _step_fused_tp(self):
        """row-sharded decode step with the glue fused into the GEMV prologues: per layer 3 fire-and-forget senders (this rank's
        slice as fp16 + epoch "LL" entries into every rank's receive buffer over NVLink peer memory, qp_xchg_send_ll) + 4 GEMVs
        (3 of them polling the receive buffer and computing residual / RMSNorm / sign / Hadamard of the gathered vector in all of
        their CTAs) + attention + the exchanging SiLU*mul/Hadamard kernel.
        Replaces the three single-CTA 8192-point norm/Hadamard kernels of the un-fused sharded list (~7 us each).
        Exchange sites and accumulator-clearing duties are those of _step_unfused (clear-before-flag protocol)."""
        L, st = lib(), torch.cuda.current_stream().cuda_stream
        sh, H, I, world, rank = self.shape, self.H, self.I, self.world, self.rank
        n0 = _cabi.launch_count()
        p = lambda t: t.data_ptr() if t is not None else None
        s_h, s_i, S = 1.0 / (math.sqrt(H) * 64.0), 1.0 / (math.sqrt(I) * 64.0), 64.0
        Ho = H // world
        keep = self._xchg_keep = []
        hc, ho = self.h, self.h2
        check(L.qp_embed(p(hc), p(self.embed), p(self.token), H, st))

        ll_site = {"attn": 4 * self.L, "acc_o": 4 * self.L + 1, "acc_dn": 4 * self.L + 2}

        def send(name, src_slice, is_f32, zero):
            """publish this rank's H / world elements of a gathered vector into every rank's LL receive buffer; returns the
            (buffer, epoch word) pair the consumer's prologue polls.  One epoch counter per BUFFER (not per layer): every use of
            the buffer must carry a value its previous contents cannot have."""
            site = ll_site[name]
            xc = self.region.xchg("ll_" + name, site)
            keep.append(xc)
            check(L.qp_xchg_send_ll(src_slice, is_f32, Ho, p(zero), zero.numel() if zero is not None else 0, ctypes.byref(xc), st))
            return self.region.base + self.region.offsets["ll_" + name][0], self.region.epoch.data_ptr() + 4 * site

        def xp(src, h_out=None, acc=None, ws=None, norm=None, su=None, scale=s_h, ll=None, ll_kind=0):
            return _cabi.XProd(p(src), p(h_out), p(acc), p(ws), S, p(norm), sh.rms_norm_eps, p(su), scale, None, None, 0, None, 0,
                               ll[0] if ll else None, ll[1] if ll else None, ll_kind)

        def run_group(projs, acc_buf, acc_off, prod):
            lead = next(i for i, (pr, _) in enumerate(projs) if pr.can_fuse())
            if len(projs) > 1:
                prod.x_out_f16 = p(self.x_h)
            projs[lead][0].launch_fused(p(acc_buf) + 4 * (acc_off + projs[lead][1]), prod, st)
            for i, (proj, off) in enumerate(projs):
                if i != lead:
                    proj.launch(p(acc_buf) + 4 * (acc_off + off), p(self.x_h), st)

        prev = None
        for li, ly in enumerate(self.layers):
            if prev is None:
                prod = xp(hc, norm=ly["norm1"], su=ly["SU_qkv"])  # acc_qkv was cleared by the previous step's last exchange
            else:
                ll = send("acc_dn", p(self.acc_dn) + 4 * rank * Ho, 1, self.acc_qkv)
                prod = xp(hc, h_out=ho, ws=prev["W_dp_full"], norm=ly["norm1"], su=ly["SU_qkv"], ll=ll, ll_kind=1)
            run_group(ly["qkv"], self.acc_qkv, 0, prod)
            if prev is not None:
                hc, ho = ho, hc
            check(L.qp_rope_attention(p(self.attn) + 2 * rank * (H // world), p(self.acc_qkv), p(ly["W_qkv"]), S, p(self.inv_freq),
                                      p(ly["kc"]), p(ly["vc"]), p(self.pos), sh.num_attention_heads // world,
                                      sh.num_key_value_heads // world, sh.head_dim, self.max_seq, ly.get("qvk", 0), None, 0, p(self.attn_scratch), st))
            ll = send("attn", p(self.attn) + 2 * rank * Ho, 0, self.acc_o)
            run_group([(ly["o"], 0)], self.acc_o, rank * Ho, xp(self.attn, su=ly["SU_o"], ll=ll, ll_kind=2))
            ll = send("acc_o", p(self.acc_o) + 4 * rank * Ho, 1, self.acc_ug)
            run_group(ly["ug"], self.acc_ug, 0, xp(hc, h_out=ho, ws=ly["W_o_full"], norm=ly["norm2"], su=ly["SU_ug"], ll=ll, ll_kind=1))
            hc, ho = ho, hc
            xc = self.region.xchg("z", 4 * li + 2)
            keep.append(xc)
            check(L.qp_silu_mul_had_grid_xchg(p(self.x_i), p(self.acc_ug), p(ly["W_ug"]), S, p(ly["SU_dp"]), I, s_i,
                                              p(self.acc_dn), self.acc_dn.numel(), p(self.sync), ctypes.byref(xc), st))
            ly["down"].launch(p(self.acc_dn) + 4 * rank * Ho, p(self.x_i), st)
            prev = ly
        xc = self.region.xchg("acc_dn", 4 * (len(self.layers) - 1) + 3)
        keep.append(xc)
        check(L.qp_fused_norm_had_xchg(p(self.xf), p(hc), 1, p(self.acc_dn), p(prev["W_dp_full"]), S, p(self.final_norm),
                                       sh.rms_norm_eps, None, H, 1.0, 0, p(self.acc_qkv), self.acc_qkv.numel(),
                                       ctypes.byref(xc), st))
        check(L.qp_gemv_f16(p(self.logits), p(self.lm_head), p(self.xf), sh.vocab_size, H, st))
        check(L.qp_argmax(p(self.token), p(self.logits), sh.vocab_size, p(self.scratch), st))
        check(L.qp_step_advance(p(self.pos), p(self.history), p(self.token), self.max_seq, st))
        self.launches_per_step = _cabi.launch_count() - n0

    def _step_fused(self):
        """one decode step with residual/RMSNorm/sign/Hadamard fused into the GEMV prologues: per layer 4 GEMVs +
        attention + SiLU*mul/Hadamard = 6 launches (9 unfused)."""
        L, st = lib(), torch.cuda.current_stream().cuda_stream
        sh, H, I = self.shape, self.H, self.I
        n0 = _cabi.launch_count()
        p = lambda t: t.data_ptr() if t is not None else None
        s_h, s_i, S = 1.0 / (math.sqrt(H) * 64.0), 1.0 / (math.sqrt(I) * 64.0), 64.0
        Hq, Hk = H, self.kvd
        hc, ho = self.h, self.h2
        check(L.qp_embed(p(hc), p(self.embed), p(self.token), H, st))

        def xp(src, h_out=None, acc=None, ws=None, norm=None, su=None, scale=s_h, x_out=None, z1=None, z2=None):
            return _cabi.XProd(p(src), p(h_out), p(acc), p(ws), S, p(norm), sh.rms_norm_eps, p(su), scale, p(x_out),
                               p(z1), z1.numel() if z1 is not None else 0, p(z2), z2.numel() if z2 is not None else 0)

        def run_group(projs, acc_buf, prod, x_buf):
            """one projection computes x in its prologue (and publishes it if siblings need it); launch order inside a
            group is free because the members write disjoint slices of the accumulator"""
            lead = next(i for i, (pr, _) in enumerate(projs) if pr.can_fuse())
            if len(projs) > 1:
                prod.x_out_f16 = p(x_buf)
            projs[lead][0].launch_fused(p(acc_buf) + 4 * projs[lead][1], prod, st)
            for i, (proj, off) in enumerate(projs):
                if i != lead:
                    proj.launch(p(acc_buf) + 4 * off, p(x_buf), st)

        prev = None
        for ly in self.layers:
            if prev is None:
                prod = xp(hc, norm=ly["norm1"], su=ly["SU_qkv"], z1=self.acc_o, z2=self.acc_ug)
            else:
                prod = xp(hc, h_out=ho, acc=self.acc_dn, ws=prev["W_dp_full"], norm=ly["norm1"], su=ly["SU_qkv"],
                          z1=self.acc_o, z2=self.acc_ug)
            run_group(ly["qkv"], self.acc_qkv, prod, self.x_h)
            if prev is not None:
                hc, ho = ho, hc
            check(L.qp_rope_attention(p(self.attn), p(self.acc_qkv), p(ly["W_qkv"]), S, p(self.inv_freq), p(ly["kc"]),
                                      p(ly["vc"]), p(self.pos), sh.num_attention_heads, sh.num_key_value_heads,
                                      sh.head_dim, self.max_seq, ly.get("qvk", 0), None, 0, p(self.attn_scratch), st))
            # fused launches clear accumulators BEFORE their dependency wait: only buffers the preceding launch does not
            # touch (the attention kernel still reads acc_qkv while the o projection starts, so ug clears it instead)
            prod = xp(self.attn, su=ly["SU_o"], z2=self.acc_dn)
            run_group([(ly["o"], 0)], self.acc_o, prod, self.x_h)
            prod = xp(hc, h_out=ho, acc=self.acc_o, ws=ly["W_o_full"], norm=ly["norm2"], su=ly["SU_ug"], z1=self.acc_qkv)
            run_group(ly["ug"], self.acc_ug, prod, self.x_h)
            hc, ho = ho, hc
            if self.silu_grid:  # one thread-block cluster, blocks exchanged through distributed shared memory
                check(L.qp_silu_mul_had_cluster(p(self.x_i), p(self.acc_ug), p(ly["W_ug"]), S, p(ly["SU_dp"]), I, s_i, None, 0,
                                                st))
            else:
                check(L.qp_silu_mul_had(p(self.x_i), p(self.acc_ug), p(ly["W_ug"]), S, p(ly["SU_dp"]), I, s_i, None, 0, st))
            ly["down"].launch(p(self.acc_dn), p(self.x_i), st)
            prev = ly
        check(L.qp_fused_norm_had(p(self.xf), p(hc), 1, p(self.acc_dn), p(prev["W_dp_full"]), S, p(self.final_norm),
                                  sh.rms_norm_eps, None, H, 1.0, 0, None, 0, st))
        check(L.qp_gemv_f16(p(self.logits), p(self.lm_head), p(self.xf), sh.vocab_size, H, st))
        check(L.qp_argmax(p(self.token), p(self.logits), sh.vocab_size, p(self.scratch), st))
        check(L.qp_step_advance(p(self.pos), p(self.history), p(self.token), self.max_seq, st))
        if hc is not self.h:  # an odd number of residual updates leaves the stream in h2: next step's embed targets self.h
            pass
        self.launches_per_step = _cabi.launch_count() - n0

    def _step_unfused(self):
        """enqueue one decode step on the current stream (eager; also what gets captured into the graph)."""
        L, st = lib(), torch.cuda.current_stream().cuda_stream
        sh, H, I = self.shape, self.H, self.I
        world, rank = self.world, self.rank
        n0 = _cabi.launch_count()
        p = lambda t: t.data_ptr() if t is not None else None
        s_h, s_i, S = 1.0 / (math.sqrt(H) * 64.0), 1.0 / (math.sqrt(I) * 64.0), 64.0
        Hq, Hk, Il, Ho = H // world, self.kvd // world, I // world, H // world
        check(L.qp_embed(p(self.h), p(self.embed), p(self.token), H, st))
        p2p, gather = self.p2p, world > 1 and not self.p2p
        keep = self._xchg_keep = []

        def norm_had(site, name, *args):
            """qp_fused_norm_had, preceded (p2p) by the in-kernel NVLink all-gather of region buffer `name`"""
            if p2p and name is not None:
                xc = self.region.xchg(name, site)
                keep.append(xc)
                check(L.qp_fused_norm_had_xchg(*args, ctypes.byref(xc), st))
            else:
                check(L.qp_fused_norm_had(*args, st))

        prev = None
        for li, ly in enumerate(self.layers):
            # residual from the previous layer's down_proj, input_layernorm, SU, Hadamard -> x_h ; zero acc_qkv
            if prev is None:
                norm_had(0, None, p(self.x_h), p(self.h), 0, None, None, 0.0, p(ly["norm1"]), sh.rms_norm_eps,
                         p(ly["SU_qkv"]), H, s_h, 1, p(self.acc_qkv), self.acc_qkv.numel())
            else:
                norm_had(4 * (li - 1) + 3, "acc_dn", p(self.x_h), p(self.h), 1, p(self.acc_dn), p(prev["W_dp_full"]), S,
                         p(ly["norm1"]), sh.rms_norm_eps, p(ly["SU_qkv"]), H, s_h, 1, p(self.acc_qkv), self.acc_qkv.numel())
            for proj, off in ly["qkv"]:
                proj.launch(p(self.acc_qkv) + 4 * off, p(self.x_h), st)
            # sharded: every rank writes its heads into its slice of the full-width buffer
            attn_dst = p(self.attn) + 2 * rank * Hq if p2p else (p(self.attn) if world == 1 else p(self.attn_loc))
            check(L.qp_rope_attention(attn_dst, p(self.acc_qkv), p(ly["W_qkv"]), S, p(self.inv_freq), p(ly["kc"]),
                                      p(ly["vc"]), p(self.pos), sh.num_attention_heads // world,
                                      sh.num_key_value_heads // world, sh.head_dim, self.max_seq, ly.get("qvk", 0), None, 0, p(self.attn_scratch), st))
            if gather:
                torch.distributed.all_gather_into_tensor(self.attn, self.attn_loc, group=self.pg)
            norm_had(4 * li + 0, "attn", p(self.x_h), p(self.attn), 0, None, None, 0.0, None, 0.0, p(ly["SU_o"]), H, s_h, 1,
                     p(self.acc_o), self.acc_o.numel())
            ly["o"].launch(p(self.acc_o) + 4 * rank * Ho, p(self.x_h), st)
            if gather:
                torch.distributed.all_gather_into_tensor(self.acc_o, self.acc_o[rank * Ho:(rank + 1) * Ho], group=self.pg)
            norm_had(4 * li + 1, "acc_o", p(self.x_h), p(self.h), 1, p(self.acc_o), p(ly["W_o_full"]), S, p(ly["norm2"]),
                     sh.rms_norm_eps, p(ly["SU_ug"]), H, s_h, 1, p(self.acc_ug), self.acc_ug.numel())
            for proj, off in ly["ug"]:
                proj.launch(p(self.acc_ug) + 4 * off, p(self.x_h), st)
            if world == 1:
                check(L.qp_silu_mul_had(p(self.x_i), p(self.acc_ug), p(ly["W_ug"]), S, p(ly["SU_dp"]), I, s_i,
                                        p(self.acc_dn), self.acc_dn.numel(), st))
            else:
                if p2p and self.silu_grid_tp:
                    # SiLU*mul + NVLink exchange of the transformed 512-blocks + cross-block Hadamard factor in one launch
                    xc = self.region.xchg("z", 4 * li + 2)
                    keep.append(xc)
                    check(L.qp_silu_mul_had_grid_xchg(p(self.x_i), p(self.acc_ug), p(ly["W_ug"]), S, p(ly["SU_dp"]), I, s_i,
                                                      p(self.acc_dn), self.acc_dn.numel(), p(self.sync), ctypes.byref(xc), st))
                else:
                    act_dst = p(self.act) + 2 * rank * Il if p2p else p(self.act_loc)
                    check(L.qp_scale_epilogue(act_dst, p(self.acc_ug), p(ly["W_ug"]), 1, 2 * Il, S, _cabi.EPI_SILU_MUL, st))
                    if gather:
                        torch.distributed.all_gather_into_tensor(self.act, self.act_loc, group=self.pg)
                    norm_had(4 * li + 2, "act", p(self.x_i), p(self.act), 0, None, None, 0.0, None, 0.0, p(ly["SU_dp"]), I,
                             s_i, 1, p(self.acc_dn), self.acc_dn.numel())
            ly["down"].launch(p(self.acc_dn) + 4 * rank * Ho, p(self.x_i), st)
            if gather:
                torch.distributed.all_gather_into_tensor(self.acc_dn, self.acc_dn[rank * Ho:(rank + 1) * Ho], group=self.pg)
            prev = ly
        norm_had(4 * (len(self.layers) - 1) + 3, "acc_dn" if world > 1 else None, p(self.xf), p(self.h), 1, p(self.acc_dn),
                 p(prev["W_dp_full"]), S, p(self.final_norm), sh.rms_norm_eps, None, H, 1.0, 0, None, 0)
        check(L.qp_gemv_f16(p(self.logits), p(self.lm_head), p(self.xf), sh.vocab_size, H, st))
        check(L.qp_argmax(p(self.token), p(self.logits), sh.vocab_size, p(self.scratch), st))
        check(L.qp_step_advance(p(self.pos), p(self.history), p(self.token), self.max_seq, st))
        self.launches_per_step = _cabi.launch_count() - n0

    def _prepare_full_scales(self):
        """Wscale vectors of o/down as full-width tensors (every rank needs all rows after the all-gather)."""
        for ly in self.layers:
            for name in ("W_o", "W_dp"):
                if self.world == 1:
                    ly[name + "_full"] = ly[name]
                else:
                    full = torch.empty(self.H, dtype=torch.float16, device=self.dev)
                    torch.distributed.all_gather_into_tensor(full, ly[name], group=self.pg)
                    ly[name + "_full"] = full

    def reset(self, token=1, pos=0):
        """restart decoding from `token` at position `pos` (pos > 0: the KV rows below pos are taken as they are in the cache)"""
        if not 0 <= pos < self.max_seq:
            raise RuntimeError(f"position {pos} outside the KV cache (max_seq={self.max_seq})")
        self.token.fill_(token)
        self.pos.fill_(pos)
        self.steps_done = pos

    def capture(self):
        self.reset()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()
        self.reset()
        return self

    def step(self):
        """one decode token.  Raises once the KV cache is full: position max_seq would be written past the cache rows and
        the attention kernel's score array (the kernel itself only clamps, to stay memory-safe)."""
        if self.steps_done >= self.max_seq:
            raise RuntimeError(f"decode position {self.steps_done} exceeds max_seq={self.max_seq}: build the DecodeRunner "
                               "with a larger max_seq or call reset()")
        self.steps_done += 1
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()

    def generate(self, n_tokens, token=1):
        """greedy decode n_tokens from `token`; returns the generated ids (host list)."""
        if n_tokens > self.max_seq:
            raise RuntimeError(f"generate({n_tokens}) exceeds max_seq={self.max_seq}")
        self.reset(token)
        for _ in range(n_tokens):
            self.step()
        torch.cuda.synchronize()
        return self.history[:n_tokens].tolist()

    def bytes_per_token(self):
        return self.weight_bytes + self.lm_head_bytes
