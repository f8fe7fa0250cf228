"""Typed Python wrappers over the C-ABI kernels (shape checks + pointer marshalling only).

Every function launches hand-written sm_100a kernels from libb200ldm.so on torch's current CUDA
stream; tensors are borrowed, outputs are caller-allocated.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream

Tensor = torch.Tensor
NUM_SMS = 148
# 2-CTA (cta_group::2) GEMM tiles.  Correct and tested, but on the AudioLDM-S shapes (one 128 x block_n tile per CTA,
# K <= 11.5k) the longer prologue / cluster syncs eat the mainloop gain (5.61 vs 5.57 ms per step): opt-in.
CTA_PAIR = os.environ.get("B200_CTA_PAIR", "1") != "0"      # 2-CTA (cta_group::2) tiles where a layer qualifies
PAIR_MIN_BN = int(os.environ.get("B200_PAIR_MIN_BN", "128"))     # 2-CTA tiles only for tiles at least this wide
TMA_BYTES_PER_CLK = float(os.environ.get("B200_TMA_BPC", "80"))    # measured: profiles/r01_gemm_mainloop_timeline.md


@dataclass
class PackedWeight:
    """bf16 [n_pad, K] K-major weight for b200_conv_gemm plus its epilogue vectors."""
    w: Tensor
    bias: Optional[Tensor]        # fp32 [n_pad] (tile-interleaved like w when geglu)
    n_valid: int                  # output columns actually stored
    block_n: int
    ntaps: int
    c0: int
    c1: int = 0
    c2: int = 0
    geglu: bool = False
    ksplit: int = 1               # > 1: split-K over this many CTAs per tile (small-M, long-K layers)
    alg_macs_per_row: Optional[float] = None   # algorithmic (unpadded) N*K of the layer; default: stored N*K
    pair: Optional[bool] = None   # 2-CTA tiles for this layer (None: the global switch decides)

    @property
    def macs_per_row(self) -> float:
        if self.alg_macs_per_row is not None:
            return self.alg_macs_per_row
        return float(self.n_valid * (2 if self.geglu else 1) * self.k)

    @property
    def n_pad(self) -> int:
        return self.w.shape[0]

    @property
    def k(self) -> int:
        return self.w.shape[1]


def box_rows(h: int, w: int, nb: int) -> Tuple[int, int]:
    """Mirror of pick_box() in conv_gemm.cu: (rows of H, images) covered by one 128-pixel tile."""
    best, out = None, (1, 1)
    bh = 1
    while bh * w <= 128:
        bni = 128 // (w * bh)
        if bni <= 256:
            tiles = math.ceil(h / bh) * math.ceil(nb / bni)
            if best is None or tiles < best or (tiles == best and bni == 1):
                best, out = tiles, (bh, bni)
        bh *= 2
    return out


def num_m_tiles(nb: int, h: int, w: int) -> int:
    bh, bni = box_rows(h, w, nb)
    return math.ceil(h / bh) * math.ceil(nb / bni)


def choose_block_n(n: int, m_tiles: int, geglu: bool = False) -> int:
    """Tile width minimising (waves x per-tile cost) on 148 SMs; n is padded up to a multiple of it."""
    best, best_cost = 64, None
    step = 128 if geglu else 64                 # the TMA-store epilogue works on 64-column output chunks
    for bn in range(step, 257, step):
        n_tiles = math.ceil(n / bn)
        waves = math.ceil(m_tiles * n_tiles / NUM_SMS)
        # per-tile time ~ MMA time (prop. to bn, floor at 64 columns: smem-read bound) + fixed overhead
        cost = waves * (max(bn, 64) + 24) * (1.0 + 0.02 * (n_tiles * bn - n) / max(n, 1))
        if best_cost is None or cost < best_cost - 1e-9 or (abs(cost - best_cost) < 1e-9 and bn > best):
            best, best_cost = bn, cost
    return best


def _kb_cycles(bn: int, pair: bool = False) -> float:
    """Cycles one 64-deep K block costs a CTA: tcgen05 time (2*bn) or operand fetch (bytes / TMA_BYTES_PER_CLK, the
    per-SM TMA ingest rate measured with the CTA-0 timeline: independent of ring depth and of how many CTAs run),
    whichever is larger.  In 2-CTA mode a CTA
    fetches its 128 A rows and only half of the weight tile."""
    return max(2.0 * bn, (16384 + bn * (64 if pair else 128)) / TMA_BYTES_PER_CLK)


# Measured per-launch model of conv_gemm_kernel (B200, isolated L2-warm replays of level-0 / level-1 / level-2 layer shapes,
# gpurun_out/r02_conv_ab.log): time = tiles per CTA x k-blocks x KB_CLK + FIXED_CLK, in SM cycles.  The fixed part is
# prologue + first operand fill + the LAST tile's epilogue (all of it exposed) + store drain + exit; 2-CTA tiles stream
# operands a little faster per tile and pay ~4 k cycles more per launch (cluster syncs, pair tail).
_KB_CLK = {(64, False): 300.0, (128, False): 350.0, (192, False): 480.0, (256, False): 620.0, (128, True): 325.0, (256, True): 620.0}
_FIXED_CLK = {(64, False): 8600.0, (128, False): 10000.0, (192, False): 12200.0, (256, False): 14400.0, (128, True): 13900.0,
              (256, True): 16700.0}
_GEGLU_EPI_CLK = 6100.0          # GEGLU epilogue per tile (erf-GELU per gate): it, not the k-loop, paces short-K tiles
TILING_MODEL = os.environ.get("B200_TILING_MODEL", "1")      # "1": the analytic model (pairs by global switch) -- measured best IN THE
                                                             # STEP (4.17 vs 4.19 ms); "2": the table fitted to isolated replays below


def choose_tiling_ex(n: int, m_tiles: int, num_kb: int, geglu: bool = False, allow_split: bool = True,
                     num_sms: int = NUM_SMS, max_bn: int = 256) -> Tuple[int, int, Optional[bool]]:
    """(block_n, ksplit, pair) minimising the measured per-launch model on `num_sms` SMs; pair None = leave the 2-CTA
    decision to the global switch (model 1).  Split-K only when it wins by > 25 %."""
    bn1, ks1 = _choose_tiling_v1(n, m_tiles, num_kb, geglu, allow_split, num_sms, max_bn)
    if TILING_MODEL == "1":
        return bn1, ks1, None
    if geglu or ks1 > 1:
        # GEGLU tiles are paced by their epilogue and split-K layers by the reduce launch: not what the table below was
        # measured on; the analytic choice stands there (measured in the step: DESIGN.md section 8)
        return bn1, ks1, None
    step = 128 if geglu else 64
    best = {}
    for ks in (1, 2, 3, 4):
        if ks > 1 and (not allow_split or geglu or num_kb < 16 * ks // 2):
            continue
        for bn in range(step, max_bn + 1, step):
            for pair in (False, True):
                if pair and not (CTA_PAIR and bn % 128 == 0 and m_tiles >= 2 and ks == 1):
                    continue
                n_tiles = math.ceil(n / bn)
                if pair:
                    tpc = math.ceil(math.ceil(m_tiles / 2) * n_tiles / max(1, num_sms // 2))
                else:
                    tiles1 = m_tiles * n_tiles
                    if ks > 1 and tiles1 > num_sms // 2:
                        continue
                    tpc = math.ceil(tiles1 * ks / num_sms)
                per_tile = math.ceil(num_kb / ks) * _KB_CLK[(bn, pair)]
                if geglu:
                    per_tile = max(per_tile, _GEGLU_EPI_CLK)
                cost = tpc * per_tile + _FIXED_CLK[(bn, pair)] + (7000 if ks > 1 else 0)
                cost *= 1.0 + 0.02 * (n_tiles * bn - n) / max(n, 1)          # padded columns are wasted stores
                key = 1 if ks == 1 else 2
                if key not in best or cost < best[key][0] - 1e-9:
                    best[key] = (cost, bn, ks, pair)
    if 2 in best and best[2][0] < 0.75 * best[1][0]:
        return best[2][1], best[2][2], False
    return best[1][1], 1, best[1][3]


def choose_tiling(n: int, m_tiles: int, num_kb: int, geglu: bool = False, allow_split: bool = True,
                  num_sms: int = NUM_SMS, max_bn: int = 256) -> Tuple[int, int]:
    """(block_n, ksplit) of choose_tiling_ex."""
    return choose_tiling_ex(n, m_tiles, num_kb, geglu, allow_split, num_sms, max_bn)[:2]


def _choose_tiling_v1(n: int, m_tiles: int, num_kb: int, geglu: bool = False, allow_split: bool = True,
                      num_sms: int = NUM_SMS, max_bn: int = 256) -> Tuple[int, int]:
    """Round-1 analytic model: a wave/cycle model with the per-SM ingest rate; split-K only when it wins by > 25 %."""
    step = 128 if geglu else 64
    best = {}
    for ks in (1, 2, 3, 4):
        if ks > 1 and (not allow_split or geglu or num_kb < 16 * ks // 2):
            continue
        for bn in range(step, max_bn + 1, step):
            pair = CTA_PAIR and bn % 128 == 0 and bn >= PAIR_MIN_BN and m_tiles >= 2 and num_kb >= 16
            tiles1 = (2 * math.ceil(m_tiles / 2) if pair else m_tiles) * math.ceil(n / bn)
            if ks > 1 and tiles1 > num_sms // 2:
                continue
            waves = math.ceil(tiles1 * ks / num_sms)
            cost = waves * (math.ceil(num_kb / ks) * _kb_cycles(bn, pair) + 2 * bn) + 4000 + (7000 if ks > 1 else 0)
            key = 1 if ks == 1 else 2
            if key not in best or cost < best[key][0] - 1e-9:
                best[key] = (cost, bn, ks)
    if 2 in best and best[2][0] < 0.75 * best[1][0]:
        return best[2][1], best[2][2]
    return best[1][1], 1


def conv_gemm(pw: PackedWeight, a0: Tensor, nb: int, h: int, w: int, out: Tensor, *, a1: Optional[Tensor] = None,
              a2: Optional[Tensor] = None, stride: int = 1, rowvec: Optional[Tensor] = None, rowvec_ld: int = 0,
              residual: Optional[Tensor] = None, out_ld: Optional[int] = None, max_ctas: int = 0,
              workspace: Optional[Tensor] = None, cta_pair: Optional[bool] = None,
              gn_stat: Optional[Tensor] = None) -> Tensor:
    """out[pix, :n_valid] = epilogue(implicit GEMM); see include/b200ldm.h::b200_conv_gemm.
    gn_stat (fp32 [nb, gn_stat_slabs(nb, h, w), n_valid / 4, 2]): also leave the partial GroupNorm statistics of the
    output for groupnorm_apply (b200_conv_gemm_gnstat)."""
    assert a0.dtype == torch.bfloat16 and a0.is_contiguous()
    assert a0.numel() == nb * h * w * pw.c0, (a0.shape, nb, h, w, pw.c0)
    if pw.c1:
        assert a1 is not None and a1.numel() == nb * h * w * pw.c1 and a1.dtype == torch.bfloat16
    if pw.c2:
        assert a2 is not None and a2.numel() == nb * h * w * pw.c2 and a2.dtype == torch.bfloat16
    out_fp32 = out.dtype == torch.float32
    assert out_fp32 or out.dtype == torch.bfloat16
    ld = out_ld if out_ld is not None else pw.n_valid
    res_ld = 0
    if residual is not None:
        assert residual.dtype == torch.bfloat16
        res_ld = pw.n_valid
    ksplit = pw.ksplit if workspace is not None else 1
    if ksplit > 1:
        assert workspace.dtype == torch.float32 and workspace.numel() >= ksplit * nb * h * w * pw.n_pad
    info = None
    if _lib.PROFILE is not None:
        m_out = nb * h * w if stride == 1 else nb * ((h - 1) // 2 + 1) * ((w - 1) // 2 + 1)
        info = {"flops": 2.0 * m_out * pw.macs_per_row, "m": nb * h * w, "n": pw.n_valid,
                "k": pw.k, "bn": pw.block_n, "taps": pw.ntaps, "desc": f"ks{ksplit}" if ksplit > 1 else None}
    if cta_pair is None and pw.pair is not None:
        cta_pair = pw.pair
    pair = (int(CTA_PAIR and pw.block_n >= PAIR_MIN_BN) if cta_pair is None else (2 if cta_pair else 0))
    if gn_stat is not None:
        assert stride == 1 and not out_fp32 and ksplit == 1 and not pw.geglu and gn_stat.dtype == torch.float32
        assert gn_stat.numel() == nb * gn_stat_slabs(nb, h, w) * (pw.n_valid // 4) * 2 and gn_stat.numel() > 0
        if info is not None:
            info["desc"] = "gnstat"
        call("b200_conv_gemm_gnstat", ptr(a0), pw.c0, ptr(a1) if pw.c1 else None, pw.c1, ptr(a2) if pw.c2 else None, pw.c2,
             nb, h, w, pw.ntaps, ptr(pw.w), pw.n_pad, pw.n_valid, ptr(pw.bias), ptr(rowvec), rowvec_ld, ptr(residual), res_ld,
             ptr(out), ld, pw.block_n, max_ctas, pair, ptr(gn_stat), stream(), info=info)
        return out
    call("b200_conv_gemm", ptr(a0), pw.c0, ptr(a1) if pw.c1 else None, pw.c1, ptr(a2) if pw.c2 else None, pw.c2,
         nb, h, w, pw.ntaps, stride, ptr(pw.w), pw.n_pad, pw.n_valid, ptr(pw.bias), ptr(rowvec), rowvec_ld,
         ptr(residual), res_ld, ptr(out), ld, int(out_fp32), int(pw.geglu), pw.block_n, max_ctas, ksplit,
         ptr(workspace) if ksplit > 1 else None, pair, stream(), info=info)
    return out


# GroupNorm statistics from the producing conv's epilogue + a one-pass apply kernel.  Parity-tested, but measured on B200
# it does not pay: the transposed read-back costs the producer ~1.2 us on its critical path, about
# what the consumer saves.  Opt-in.
def gemm_nt(a: Tensor, b: Tensor, n_valid: int, out: Tensor, block_n: int, out_ld: Optional[int] = None) -> Tensor:
    """out[m, :n_valid] = a[m, k] . b[:n_valid, k]^T, both bf16 K-major activations written earlier in the stream (see
    include/b200ldm.h::b200_gemm_nt).  b: [b_rows, k] with b_rows a multiple of block_n (rows >= n_valid: don't-care)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous()
    m, k = a.shape
    assert b.shape[1] == k and b.shape[0] % block_n == 0 and b.shape[0] >= n_valid and k % 64 == 0 and n_valid % 8 == 0
    out_fp32 = out.dtype == torch.float32
    ld = out_ld if out_ld is not None else n_valid
    info = {"flops": 2.0 * m * n_valid * k, "m": m, "n": n_valid, "k": k, "bn": block_n, "taps": 1, "desc": "nt"} \
        if _lib.PROFILE is not None else None
    call("b200_gemm_nt", ptr(a), m, k, ptr(b), b.shape[0], n_valid, ptr(out), ld, int(out_fp32), block_n,
         int(CTA_PAIR and block_n >= PAIR_MIN_BN), stream(), info=info)
    return out


GN_ONEPASS = os.environ.get("B200_GN_ONEPASS", "0") != "0"
_SLABS: dict = {}


def gn_stat_slabs(nb: int, h: int, w: int) -> int:
    """Slabs per image of the GroupNorm partial statistics a conv over [nb, h, w, *] leaves (0: unsupported geometry)."""
    key = (nb, h, w)
    if key not in _SLABS:
        _SLABS[key] = int(_lib.load().b200_gn_stat_slabs(nb, h, w))
    return _SLABS[key]


def groupnorm_apply(x0: Tensor, c0: int, st0: Tensor, x1: Optional[Tensor], c1: int, st1: Optional[Tensor], nb: int, hw: int,
                    gamma: Tensor, beta: Tensor, eps: float, silu: bool, y: Tensor, groups: int = 32) -> Tensor:
    """One-pass GroupNorm (+SiLU) over cat(x0, x1) from the producers' partial statistics; see b200_groupnorm_apply."""
    assert x0.dtype == torch.bfloat16 and y.dtype == torch.bfloat16 and st0.dtype == torch.float32
    assert gamma.numel() == c0 + c1 and gamma.dtype == torch.float32
    slabs0 = st0.numel() // (nb * (c0 // 4) * 2)
    slabs1 = st1.numel() // (nb * (c1 // 4) * 2) if c1 else 0
    info = {"desc": f"nb{nb} hw{hw} c{c0}+{c1}", "bytes": 4.0 * nb * hw * (c0 + c1)} if _lib.PROFILE is not None else None
    call("b200_groupnorm_apply", ptr(x0), c0, ptr(st0), slabs0, ptr(x1) if c1 else None, c1, ptr(st1) if c1 else None, slabs1,
         nb, hw, groups, ptr(gamma), ptr(beta), float(eps), int(silu), ptr(y), stream(), info=info)
    return y


LORA_FUSED = os.environ.get("B200_LORA_FUSED", "1") != "0"
LORA_FUSED_MAX_BN = 192          # 2 * block_n accumulator columns + up to 64 columns of T must fit 512 TMEM columns


def lora_fusion_pays(n: int, m_tiles: int, kp: int, num_sms: int = NUM_SMS) -> bool:
    """Plan-time choice.  The in-kernel down-projection costs ~3.5 us per tile (phase 0 is on the tile's critical path)
    and caps the tile width at 192; it replaces a ~6 us launch.  Measured on B200: a win whenever a CTA runs at most
    ~2 tiles (all level-2/3 layers, to_out at level 1), a loss for the multi-wave level-1 QKV GEMM (20.8 vs 19.8 us)."""
    return LORA_FUSED and kp == 64 and m_tiles * math.ceil(n / LORA_FUSED_MAX_BN) <= 2 * num_sms


def linear_lora_ok(pw: PackedWeight, down: Optional[PackedWeight]) -> bool:
    """Can this adapted linear layer run with the LoRA down-projection inside the GEMM kernel?"""
    return (LORA_FUSED and getattr(pw, "lora_fused", False) and down is not None and getattr(down, "lora_rows", 0) > 0 and
            pw.c1 == 64 and pw.c2 == 0 and
            pw.ntaps == 1 and not pw.geglu and pw.ksplit == 1 and pw.block_n <= LORA_FUSED_MAX_BN and down.n_pad == 64)


def linear_lora(pw: PackedWeight, down: PackedWeight, x: Tensor, m: int, out: Tensor, *, residual: Optional[Tensor] = None,
                t_out: Optional[Tensor] = None, max_ctas: int = 0) -> Tensor:
    """out = x W^T + (x A^T)(s B)^T (+ bias + residual) in ONE launch; see include/b200ldm.h::b200_linear_lora."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == m * pw.c0 and out.dtype == torch.bfloat16
    assert down.w.shape == (64, pw.c0) and pw.c1 == 64
    res_ld = pw.n_valid if residual is not None else 0
    if t_out is not None:
        assert t_out.dtype == torch.bfloat16 and t_out.numel() == m * 64
    info = None
    if _lib.PROFILE is not None:
        info = {"flops": 2.0 * m * pw.macs_per_row, "m": m, "n": pw.n_valid, "k": pw.k, "bn": pw.block_n, "taps": 1,
                "desc": "lora-fused"}
    call("b200_linear_lora", ptr(x), pw.c0, m, ptr(pw.w), pw.n_pad, pw.n_valid, ptr(pw.bias), ptr(residual), res_ld,
         ptr(out), pw.n_valid, pw.block_n, max_ctas, ptr(down.w), down.lora_rows, ptr(t_out), stream(), info=info)
    return out


# LayerNorm folded into the consuming GEMM (b200_linear_ln + b200_linear_stats).  Correct and tested, removes 48 launches
# per step, but measured on B200 it does not pay yet: 5.36 vs 5.25 ms per step (the folded level-1 QKV GEMM needs
# 192-wide tiles for the in-kernel LoRA branch and loses more than the LayerNorm launch cost).  Opt-in.
LN_FUSED = os.environ.get("B200_LN_FUSED", "0") != "0"
# Selective form: fold a LayerNorm only where the layer has at most this many 128-row m-tiles (the deep UNet levels are
# launch-latency-bound, the shallow ones pay for the 192-column tile cap).  -1: no limit (everything when LN_FUSED).
LN_FUSED_QKV_MAX_MT = int(os.environ.get("B200_LN_FUSED_QKV_MT", "-1"))     # norm1 / norm2 -> the QKV GEMM
LN_FUSED_FF_MAX_MT = int(os.environ.get("B200_LN_FUSED_FF_MT", "-1"))       # norm3 -> ff.net.0.proj (GEGLU)


def ln_fusion_wanted(kind: str, m_tiles: int) -> bool:
    lim = LN_FUSED_QKV_MAX_MT if kind == "qkv" else LN_FUSED_FF_MAX_MT
    return LN_FUSED and (lim < 0 or m_tiles <= lim)


def linear_stats(pw: PackedWeight, x: Tensor, m: int, out: Tensor, stat_out: Tensor, *, a1: Optional[Tensor] = None,
                 down: Optional[PackedWeight] = None, residual: Optional[Tensor] = None, max_ctas: int = 0) -> Tensor:
    """A linear layer that also leaves its output's per-row, per-64-column (sum, sum of squares) in stat_out
    [m, n_valid / 64, 2] for a following LayerNorm-folded GEMM; see include/b200ldm.h::b200_linear_stats."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == m * pw.c0 and out.dtype == torch.bfloat16
    assert stat_out.dtype == torch.float32 and stat_out.numel() == m * (pw.n_valid // 64) * 2 and pw.n_valid % 64 == 0
    assert pw.ksplit == 1 and not pw.geglu and pw.ntaps == 1
    res_ld = pw.n_valid if residual is not None else 0
    info = None
    if _lib.PROFILE is not None:
        info = {"flops": 2.0 * m * pw.macs_per_row, "m": m, "n": pw.n_valid, "k": pw.k, "bn": pw.block_n, "taps": 1,
                "desc": "stats" + ("+lora" if down is not None else "")}
    call("b200_linear_stats", ptr(x), pw.c0, ptr(a1) if (pw.c1 and down is None) else None, pw.c1, m, ptr(pw.w), pw.n_pad,
         pw.n_valid, ptr(pw.bias), ptr(residual), res_ld, ptr(out), pw.n_valid, pw.block_n, max_ctas,
         ptr(down.w) if down is not None else None, down.lora_rows if down is not None else 0, ptr(stat_out), stream(),
         info=info)
    return out


def linear_ln(pw: PackedWeight, x: Tensor, m: int, out: Tensor, stats: Tensor, *, eps: float = 1e-5,
              down: Optional[PackedWeight] = None, residual: Optional[Tensor] = None, max_ctas: int = 0) -> Tensor:
    """out = LN(x) W^T (+ LoRA) with the LayerNorm folded into the GEMM; pw / down are `.ln`-packed (gamma folded in,
    pw.ln_g / down.ln_g / down.bias set); see include/b200ldm.h::b200_linear_ln."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == m * pw.c0 and out.dtype == torch.bfloat16
    assert pw.bias is not None and getattr(pw, "ln_g", None) is not None and pw.ksplit == 1
    assert (pw.c1 == 64) == (down is not None)
    assert stats.dtype == torch.float32 and stats.numel() == m * (pw.c0 // 64) * 2
    res_ld = pw.n_valid if residual is not None else 0
    info = None
    if _lib.PROFILE is not None:
        info = {"flops": 2.0 * m * pw.macs_per_row, "m": m, "n": pw.n_valid, "k": pw.k, "bn": pw.block_n, "taps": 1,
                "desc": "ln-fused" + ("+lora" if down is not None else "") + ("+geglu" if pw.geglu else "")}
    call("b200_linear_ln", ptr(x), pw.c0, m, ptr(pw.w), pw.n_pad, pw.n_valid, ptr(pw.bias), ptr(pw.ln_g), float(eps),
         ptr(stats), ptr(residual), res_ld, ptr(out), pw.n_valid, int(pw.geglu), pw.block_n, max_ctas,
         ptr(down.w) if down is not None else None, down.lora_rows if down is not None else 0,
         ptr(down.ln_g) if down is not None else None, ptr(down.bias) if down is not None else None, stream(), info=info)
    return out


def set_sm_budget(n: int) -> None:
    """Launch-shape hint for the following launches of this thread (0 = the whole GPU); see b200_set_sm_budget."""
    _lib.check(_lib.load().b200_set_sm_budget(int(n)), "b200_set_sm_budget")


def groupnorm_silu(x0: Tensor, c0: int, x1: Optional[Tensor], c1: int, nb: int, hw: int, gamma: Tensor, beta: Tensor,
                   eps: float, silu: bool, y: Tensor, groups: int = 32) -> Tensor:
    assert x0.dtype == torch.bfloat16 and y.dtype == torch.bfloat16
    assert gamma.numel() == c0 + c1 and gamma.dtype == torch.float32
    info = {"desc": f"nb{nb} hw{hw} c{c0}+{c1}", "bytes": 4.0 * nb * hw * (c0 + c1)} if _lib.PROFILE is not None else None
    call("b200_groupnorm_silu", ptr(x0), c0, ptr(x1) if c1 else None, c1, nb, hw, groups, ptr(gamma), ptr(beta),
         float(eps), int(silu), ptr(y), stream(), info=info)
    return y


def softmax_rows(s: Tensor, rows: int, cols: int, cols_pad: int, p: Tensor, scale: float = 1.0) -> Tensor:
    """p[r, :cols] = softmax(scale * s[r, :cols]), zero up to cols_pad; s fp32 [rows, ld_s], p bf16 [rows, ld_p]."""
    assert s.dtype == torch.float32 and p.dtype == torch.bfloat16 and s.dim() == 2 and p.dim() == 2
    assert s.stride(1) == 1 and p.stride(1) == 1 and s.shape[0] >= rows and p.shape[0] >= rows
    call("b200_softmax_rows", ptr(s), rows, cols, cols_pad, s.stride(0), ptr(p), p.stride(0), float(scale), stream())
    return p


def layernorm(x: Tensor, m: int, c: int, gamma: Tensor, beta: Tensor, eps: float, y: Tensor) -> Tensor:
    assert x.dtype == torch.bfloat16 and y.dtype == torch.bfloat16
    info = {"desc": f"m{m} c{c}", "bytes": 4.0 * m * c} if _lib.PROFILE is not None else None
    call("b200_layernorm", ptr(x), m, c, ptr(gamma), ptr(beta), float(eps), ptr(y), stream(), info=info)
    return y


def embed_layernorm(ids: Tensor, pos_ids: Tensor, word: Tensor, pos: Tensor, type0: Tensor, gamma: Tensor, beta: Tensor,
                    eps: float, y: Tensor) -> Tensor:
    """y[row] = LayerNorm(word[ids[row]] + type0 + pos[pos_ids[row]]) as bf16; see include/b200ldm.h::b200_embed_layernorm."""
    m, c = ids.numel(), word.shape[1]
    assert ids.dtype == torch.int32 and pos_ids.dtype == torch.int32 and pos_ids.numel() == m
    for t in (word, pos, type0, gamma, beta):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert y.dtype == torch.bfloat16 and y.numel() == m * c and pos.shape[1] == c and type0.numel() == c
    info = {"desc": f"m{m} c{c}", "bytes": 14.0 * m * c} if _lib.PROFILE is not None else None
    call("b200_embed_layernorm", ptr(ids), ptr(pos_ids), m, c, word.shape[0], pos.shape[0], ptr(word), ptr(pos), ptr(type0),
         ptr(gamma), ptr(beta), float(eps), ptr(y), stream(), info=info)
    return y


def attention(qkv: Tensor, out: Tensor, batch: int, seq: int, heads: int, head_dim: int,
              scale: Optional[float] = None, variant: int = 0) -> Tensor:
    assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
    assert qkv.numel() == batch * seq * 3 * heads * head_dim
    if scale is None:
        scale = head_dim ** -0.5
    info = ({"flops": 4.0 * batch * heads * seq * seq * head_dim, "desc": f"b{batch} s{seq} d{head_dim}"}
            if _lib.PROFILE is not None else None)
    call("b200_attention", ptr(qkv), ptr(out), batch, seq, heads, head_dim, float(scale), variant, stream(), info=info)
    return out


def time_class_embed(t_steps: Tensor, step_ptr: Optional[Tensor], per_sample: bool, labels: Tensor, nb: int,
                     tproj: int, ted: int, class_in: int, w1, b1, w2, b2, wc, bc, emb: Optional[Tensor],
                     silu_emb: Tensor) -> None:
    assert t_steps.dtype == torch.float32 and labels.dtype == torch.float32 and silu_emb.dtype == torch.bfloat16
    call("b200_time_class_embed", ptr(t_steps), ptr(step_ptr), int(per_sample), ptr(labels), nb, tproj, ted, class_in,
         ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(wc), ptr(bc), ptr(emb), ptr(silu_emb), stream())


def pack_nchw_to_nhwc(x: Tensor, nb: int, c: int, hw: int, c_pad: int, y: Tensor) -> Tensor:
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16 and x.is_contiguous()
    call("b200_pack_nchw_to_nhwc", ptr(x), nb, c, hw, c_pad, ptr(y), stream())
    return y


def unpack_nhwc_to_nchw(x: Tensor, nb: int, c: int, hw: int, y: Tensor) -> Tensor:
    assert x.dtype == torch.float32 and y.dtype == torch.float32
    call("b200_unpack_nhwc_to_nchw", ptr(x), nb, c, hw, ptr(y), stream())
    return y


def upsample_nearest(x: Tensor, nb: int, h: int, w: int, c: int, ho: int, wo: int, y: Tensor) -> Tensor:
    assert x.dtype == torch.bfloat16 and y.dtype == torch.bfloat16
    info = {"desc": f"nb{nb} {h}x{w}->{ho}x{wo} c{c}", "bytes": 2.0 * nb * c * (h * w + ho * wo)} if _lib.PROFILE is not None else None
    call("b200_upsample_nearest", ptr(x), nb, h, w, c, ho, wo, ptr(y), stream(), info=info)
    return y


def sampler_step(eps: Tensor, x: Tensor, x_saved: Optional[Tensor], hist: Optional[Tensor], table: Tensor,
                 step_ptr: Tensor, guidance: float, do_cfg: bool, nb: int, hw: int, c: int, c_pad: int,
                 xin_next: Optional[Tensor]) -> None:
    assert eps.dtype == torch.float32 and x.dtype == torch.float32 and table.dtype == torch.float32
    assert step_ptr.dtype == torch.int32
    # algorithmic bytes: read x, e_u, e_t (fp32), write x' (fp32) and the next CFG-duplicated bf16 UNet input
    info = {"desc": f"nb{nb} hw{hw}", "bytes": float(nb * hw * c) * ((3 if do_cfg else 2) * 4 + 4 + (2 if do_cfg else 1) * 2)} \
        if _lib.PROFILE is not None else None
    call("b200_sampler_step", ptr(eps), ptr(x), ptr(x_saved), ptr(hist), ptr(table), ptr(step_ptr), float(guidance),
         int(do_cfg), nb, hw, c, c_pad, ptr(xin_next), stream(), info=info)


def add_noise(x0: Tensor, noise: Tensor, sqrt_ac: Tensor, sqrt_1mac: Tensor, out: Tensor) -> Tensor:
    nb, c = x0.shape[0], x0.shape[1]
    hw = x0[0, 0].numel()
    call("b200_add_noise", ptr(x0), ptr(noise), ptr(sqrt_ac), ptr(sqrt_1mac), nb, c, hw, ptr(out), stream())
    return out


def adamw_flat(param: Tensor, grad: Tensor, m: Tensor, v: Tensor, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float, step: int, grad_scale: float = 1.0) -> None:
    assert all(t.dtype == torch.float32 and t.is_contiguous() for t in (param, grad, m, v))
    call("b200_adamw_flat", ptr(param), ptr(grad), ptr(m), ptr(v), param.numel(), lr, beta1, beta2, eps, weight_decay,
         step, grad_scale, stream())


def adamw_flat_dev(param: Tensor, grad: Tensor, m: Tensor, v: Tensor, hyper: Tensor, beta1: float, beta2: float,
                   eps: float, weight_decay: float) -> None:
    """hyper: device fp32 [4] = {lr, 1 - beta1^t, 1 - beta2^t, grad_scale}."""
    assert all(t.dtype == torch.float32 and t.is_contiguous() for t in (param, grad, m, v, hyper)) and hyper.numel() >= 4
    call("b200_adamw_flat_dev", ptr(param), ptr(grad), ptr(m), ptr(v), param.numel(), ptr(hyper), beta1, beta2, eps,
         weight_decay, stream())


def mse_partial(pred: Tensor, target: Tensor, out_sum: Tensor) -> None:
    call("b200_mse_partial", ptr(pred), ptr(target), pred.numel(), ptr(out_sum), stream())


# ------------------------------------------------------------------------------------------ fine-tuning step
class WgradDesc(ctypes.Structure):
    """One (U, V) pair of b200_lora_wgrad: G[c, j] += scale * sum_m U[m, c] V[m, j] -> out[c*ldc + j*ldj] (fp32)."""
    _fields_ = [("u", ctypes.c_void_p), ("v", ctypes.c_void_p), ("out", ctypes.c_void_p), ("ldu", ctypes.c_int),
                ("ldv", ctypes.c_int), ("C", ctypes.c_int), ("r", ctypes.c_int), ("ldc", ctypes.c_int),
                ("ldj", ctypes.c_int), ("scale", ctypes.c_float)]


REFRESH_DTYPE = [("dst", "<u8"), ("src_off", "<i8"), ("dst_ld", "<i4"), ("src_rows", "<i4"), ("src_cols", "<i4"),
                 ("transpose", "<i4"), ("scale", "<f4"), ("pad", "<i4")]      # struct RefreshDesc in csrc/train.cu


def attention_lse(qkv: Tensor, out: Tensor, lse: Tensor, batch: int, seq: int, heads: int, head_dim: int,
                  scale: Optional[float] = None) -> Tensor:
    assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and lse.dtype == torch.float32
    assert lse.numel() == batch * heads * seq
    if scale is None:
        scale = head_dim ** -0.5
    info = ({"flops": 4.0 * batch * heads * seq * seq * head_dim, "desc": f"b{batch} s{seq} d{head_dim}"}
            if _lib.PROFILE is not None else None)
    call("b200_attention_lse", ptr(qkv), ptr(out), ptr(lse), batch, seq, heads, head_dim, float(scale), stream(), info=info)
    return out


def attention_bwd(qkv: Tensor, o: Tensor, dout: Tensor, lse: Tensor, delta: Tensor, dqkv: Tensor, batch: int, seq: int,
                  heads: int, head_dim: int, scale: Optional[float] = None) -> Tensor:
    for t in (qkv, o, dout, dqkv):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
    assert lse.dtype == torch.float32 and delta.dtype == torch.float32 and delta.numel() == lse.numel()
    if scale is None:
        scale = head_dim ** -0.5
    info = ({"flops": 14.0 * batch * heads * seq * seq * head_dim, "desc": f"b{batch} s{seq} d{head_dim}"}
            if _lib.PROFILE is not None else None)
    call("b200_attention_bwd", ptr(qkv), ptr(o), ptr(dout), ptr(lse), ptr(delta), ptr(dqkv), batch, seq, heads, head_dim,
         float(scale), stream(), info=info)
    return dqkv


def groupnorm_silu_stats(x0: Tensor, c0: int, x1: Optional[Tensor], c1: int, nb: int, hw: int, gamma: Tensor,
                         beta: Tensor, eps: float, silu: bool, y: Tensor, stats: Tensor, groups: int = 32) -> Tensor:
    assert x0.dtype == torch.bfloat16 and y.dtype == torch.bfloat16 and stats.dtype == torch.float32
    assert stats.numel() == nb * groups * 2
    call("b200_groupnorm_silu_stats", ptr(x0), c0, ptr(x1) if c1 else None, c1, nb, hw, groups, ptr(gamma), ptr(beta),
         float(eps), int(silu), ptr(y), ptr(stats), stream())
    return y


def groupnorm_silu_bwd(x0: Tensor, c0: int, x1: Optional[Tensor], c1: int, nb: int, hw: int, gamma: Tensor, beta: Tensor,
                       stats: Tensor, silu: bool, dy: Tensor, dres: Optional[Tensor], res_ld: int, dx0: Tensor,
                       dx1: Optional[Tensor], groups: int = 32) -> None:
    assert dy.dtype == torch.bfloat16 and dx0.dtype == torch.bfloat16 and dy.numel() == nb * hw * (c0 + c1)
    info = {"desc": f"nb{nb} hw{hw} c{c0}+{c1}", "bytes": 6.0 * nb * hw * (c0 + c1)} if _lib.PROFILE is not None else None
    call("b200_groupnorm_silu_bwd", ptr(x0), c0, ptr(x1) if c1 else None, c1, nb, hw, groups, ptr(gamma), ptr(beta),
         ptr(stats), int(silu), ptr(dy), ptr(dres), res_ld, ptr(dx0), ptr(dx1), stream(), info=info)


def layernorm_bwd(x: Tensor, dy: Tensor, m: int, c: int, gamma: Tensor, eps: float, dres: Optional[Tensor],
                  dx: Tensor) -> Tensor:
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and dx.dtype == torch.bfloat16
    call("b200_layernorm_bwd", ptr(x), ptr(dy), m, c, ptr(gamma), float(eps), ptr(dres), ptr(dx), stream())
    return dx


def geglu_fwd(h: Tensor, m: int, f: int, out: Tensor) -> Tensor:
    assert h.dtype == torch.bfloat16 and h.numel() == m * 2 * f and out.numel() == m * f
    call("b200_geglu_fwd", ptr(h), m, f, ptr(out), stream())
    return out


def geglu_bwd(h: Tensor, dout: Tensor, m: int, f: int, dh: Tensor) -> Tensor:
    assert h.dtype == torch.bfloat16 and dout.numel() == m * f and dh.numel() == m * 2 * f
    call("b200_geglu_bwd", ptr(h), ptr(dout), m, f, ptr(dh), stream())
    return dh


def lora_wgrad(descs: Sequence["WgradDesc"], m: int) -> None:
    arr = (WgradDesc * len(descs))(*descs)
    call("b200_lora_wgrad", ctypes.cast(arr, ctypes.c_void_p), len(descs), m, stream())


def zero_insert(dy: Tensor, nb: int, h: int, w: int, c: int, z: Tensor) -> Tensor:
    assert dy.dtype == torch.bfloat16 and z.numel() == nb * h * w * c
    call("b200_zero_insert", ptr(dy), nb, h, w, c, ptr(z), stream())
    return z


def upsample_nearest_bwd(dy: Tensor, nb: int, h: int, w: int, c: int, ho: int, wo: int, dx: Tensor) -> Tensor:
    assert dy.dtype == torch.bfloat16 and dy.numel() == nb * ho * wo * c and dx.numel() == nb * h * w * c
    call("b200_upsample_nearest_bwd", ptr(dy), nb, h, w, c, ho, wo, ptr(dx), stream())
    return dx


def add_bf16(y: Tensor, x: Tensor) -> Tensor:
    assert y.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and y.numel() == x.numel()
    call("b200_add_bf16", ptr(y), ptr(x), y.numel(), stream())
    return y


def mse_grad(pred_nhwc: Tensor, noise_nchw: Tensor, nb: int, hw: int, c_pad: int, inv_count: float, loss_sum: Tensor,
             deps: Tensor) -> None:
    assert pred_nhwc.dtype == torch.float32 and noise_nchw.dtype == torch.float32 and deps.dtype == torch.bfloat16
    assert deps.numel() == nb * hw * c_pad and loss_sum.dtype == torch.float32
    call("b200_mse_grad", ptr(pred_nhwc), ptr(noise_nchw), nb, hw, c_pad, float(inv_count), ptr(loss_sum), ptr(deps), stream())


def lora_refresh(descs_dev: Tensor, n: int, flat: Tensor) -> None:
    assert descs_dev.dtype == torch.uint8 and flat.dtype == torch.float32
    call("b200_lora_refresh", ptr(descs_dev), n, ptr(flat), stream())


# ------------------------------------------------------------------------------------------ HiFi-GAN vocoder
def conv1d(pw: PackedWeight, x: Tensor, nb: int, length: int, out: Tensor, *, dh0: int, dh_step: int,
           m_rows: Optional[int] = None, residual: Optional[Tensor] = None, res_slope: float = 1.0, act_slope: float = 1.0,
           act_tanh: bool = False, out_ld: Optional[int] = None, out_batch_stride: int = 0,
           cta_pair: Optional[bool] = None) -> Tensor:
    """One Conv1d (or one output phase of a ConvTranspose1d) over time-major x [nb, length, c]; see
    include/b200ldm.h::b200_conv1d.  residual is stored post-LeakyReLU(res_slope) (1.0: plain).  act_tanh: False = LeakyReLU
    (act_slope; 0.0 = ReLU, 1.0 = identity), True = tanh, 2 = exact-erf GELU."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == nb * length * pw.c0 and pw.c1 == 0 and pw.c2 == 0
    rows = length if m_rows is None else m_rows
    out_fp32 = out.dtype == torch.float32
    assert out_fp32 or out.dtype == torch.bfloat16
    ld = out_ld if out_ld is not None else pw.n_valid
    res_ld = 0
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.numel() == nb * length * pw.n_valid
        res_ld = pw.n_valid
    info = None
    if _lib.PROFILE is not None:
        info = {"flops": 2.0 * nb * rows * pw.macs_per_row, "m": nb * rows, "n": pw.n_valid, "k": pw.k, "bn": pw.block_n,
                "taps": pw.ntaps, "desc": "conv1d"}
    pair = (int(CTA_PAIR and pw.block_n >= PAIR_MIN_BN) if cta_pair is None else (2 if cta_pair else 0))
    call("b200_conv1d", ptr(x), pw.c0, nb, length, pw.ntaps, dh0, dh_step, rows, ptr(pw.w), pw.n_pad, pw.n_valid, ptr(pw.bias),
         ptr(residual), res_ld, 1.0 / res_slope, ptr(out), ld, out_batch_stride, int(out_fp32), float(act_slope),
         int(act_tanh), pw.block_n, pair, stream(), info=info)
    return out


def lrelu_mean3(a0: Tensor, a1: Tensor, a2: Tensor, in_slope: float, out_slope: float, y: Tensor) -> Tensor:
    """y = leaky_relu((x0 + x1 + x2) / 3, out_slope) with a_i = leaky_relu(x_i, in_slope) stored; see b200_lrelu_mean3."""
    for t in (a0, a1, a2, y):
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == y.numel()
    info = {"desc": f"n{y.numel()}", "bytes": 8.0 * y.numel()} if _lib.PROFILE is not None else None
    call("b200_lrelu_mean3", ptr(a0), ptr(a1), ptr(a2), y.numel(), float(in_slope), float(out_slope), ptr(y), stream(), info=info)
    return y


def f32_to_bf16(x: Tensor, y: Tensor) -> Tensor:
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16 and x.is_contiguous() and y.is_contiguous()
    assert x.numel() == y.numel()
    call("b200_f32_to_bf16", ptr(x), x.numel(), ptr(y), stream())
    return y


def conv3x3_s2_pad01(pw: PackedWeight, x: Tensor, nb: int, h: int, w: int, out: Tensor, cta_pair: Optional[bool] = None) -> Tensor:
    """Downsample2D(padding=0) of the VAE encoder: F.pad(x, (0, 1, 0, 1)) + conv k3 s2; see b200_conv3x3_s2_pad01."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == nb * h * w * pw.c0 and pw.ntaps == 9
    assert out.dtype == torch.bfloat16 and pw.c1 == 0 and pw.c2 == 0 and pw.ksplit == 1
    ho, wo = (h - 2) // 2 + 1, (w - 2) // 2 + 1
    assert out.numel() == nb * ho * wo * pw.n_valid
    info = {"flops": 2.0 * nb * ho * wo * pw.macs_per_row, "m": nb * ho * wo, "n": pw.n_valid, "k": pw.k, "bn": pw.block_n, "taps": 9,
            "desc": "s2pad01"} if _lib.PROFILE is not None else None
    pair = (int(CTA_PAIR and pw.block_n >= PAIR_MIN_BN) if cta_pair is None else (2 if cta_pair else 0))
    call("b200_conv3x3_s2_pad01", ptr(x), pw.c0, nb, h, w, ptr(pw.w), pw.n_pad, pw.n_valid, ptr(pw.bias), ptr(out), pw.n_valid,
         pw.block_n, pair, stream(), info=info)
    return out
