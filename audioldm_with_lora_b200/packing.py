"""Weight packing for b200_conv_gemm: diffusers-layout fp32 tensors -> bf16 [n_pad, K] K-major.

K ordering of a 3x3 conv is tap-major (kh, kw, cin) so that one 64-channel K block is one TMA box
of the NHWC activation at tap offset (kh-1, kw-1).  Extra 1x1 K segments (conv_shortcut over the two
halves of cat([h, skip]); the LoRA up-projection s*B) are appended along K.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .ops import PackedWeight

Tensor = torch.Tensor


def _ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def conv3x3_to_k(w: Tensor, c_pad: Optional[int] = None) -> Tensor:
    """[Co, Ci, 3, 3] -> [Co, 9 * Ci_pad] with K index = (kh*3 + kw) * Ci_pad + ci."""
    co, ci = w.shape[0], w.shape[1]
    cp = c_pad or _ceil_to(ci, 64)
    out = w.new_zeros(co, 3, 3, cp)
    out[..., :ci] = w.permute(0, 2, 3, 1)
    return out.reshape(co, 9 * cp)


def pad_k(w: Tensor, k_pad: int) -> Tensor:
    if w.shape[1] == k_pad:
        return w
    out = w.new_zeros(w.shape[0], k_pad)
    out[:, : w.shape[1]] = w
    return out


def pack(segments: Sequence[Tensor], bias: Optional[Tensor], block_n: int, ntaps: int, c0: int, c1: int = 0,
         c2: int = 0, geglu: bool = False, device=None, ksplit: int = 1) -> PackedWeight:
    """segments: fp32 [N, K_i] matrices concatenated along K (already tap-major / channel-padded)."""
    w = torch.cat([s.float() for s in segments], dim=1)
    n, k = w.shape
    assert k == ntaps * c0 + c1 + c2, (k, ntaps, c0, c1, c2)
    assert c0 % 64 == 0 and c1 % 64 == 0 and c2 % 64 == 0
    b = bias.float() if bias is not None else None
    if geglu:
        # rows [0, n/2) are values, [n/2, n) gates -> per tile: [bn/2 values | bn/2 gates]
        assert n % 2 == 0 and block_n % 64 == 0
        half, hb = n // 2, block_n // 2
        n_half_pad = _ceil_to(half, hb)
        wv = w.new_zeros(n_half_pad, k); wv[:half] = w[:half]
        wg = w.new_zeros(n_half_pad, k); wg[:half] = w[half:]
        w = torch.stack([wv.view(-1, hb, k), wg.view(-1, hb, k)], dim=1).reshape(-1, k)
        if b is not None:
            bv = b.new_zeros(n_half_pad); bv[:half] = b[:half]
            bg = b.new_zeros(n_half_pad); bg[:half] = b[half:]
            b = torch.stack([bv.view(-1, hb), bg.view(-1, hb)], dim=1).reshape(-1)
        n_valid = half
    else:
        n_pad = _ceil_to(n, block_n)
        if n_pad != n:
            wp = w.new_zeros(n_pad, k); wp[:n] = w; w = wp
            if b is not None:
                bp = b.new_zeros(n_pad); bp[:n] = b; b = bp
        n_valid = n
    assert n_valid % 8 == 0
    dev = device if device is not None else w.device
    return PackedWeight(w=w.to(dev, torch.bfloat16).contiguous(),
                        bias=None if b is None else b.to(dev, torch.float32).contiguous(),
                        n_valid=n_valid, block_n=block_n, ntaps=ntaps, c0=c0, c1=c1, c2=c2, geglu=geglu, ksplit=ksplit)


def lora_pad(r_total: int) -> int:
    """K columns the LoRA segment occupies (T = x A^T is stored [M, lora_pad])."""
    return _ceil_to(max(r_total, 1), 64)


def pack_lora_down(a_list: List[Optional[Tensor]], c: int, block_n: int = 64, device=None) -> PackedWeight:
    """Stack the A matrices ([r, C] each, None = not adapted) -> [lora_pad, C]; T = x . stack^T."""
    r_tot = sum(a.shape[0] for a in a_list if a is not None)
    n = lora_pad(r_tot)
    w = torch.zeros(n, c)
    off = 0
    for a in a_list:
        if a is None:
            continue
        w[off: off + a.shape[0]] = a.float().cpu()
        off += a.shape[0]
    return pack([w], None, block_n=min(block_n, n), ntaps=1, c0=c, device=device)


def lora_up_segment(b_list: List[Optional[Tensor]], a_list: List[Optional[Tensor]], scales: List[float],
                    c_out_each: int) -> Tensor:
    """Block-diagonal [len*c_out, lora_pad] matrix whose block i is scales[i] * B_i at T's column offset of A_i."""
    r_tot = sum(a.shape[0] for a in a_list if a is not None)
    kp = lora_pad(r_tot)
    seg = torch.zeros(len(b_list) * c_out_each, kp)
    off = 0
    for i, (b, a) in enumerate(zip(b_list, a_list)):
        if a is None:
            continue
        r = a.shape[0]
        seg[i * c_out_each: (i + 1) * c_out_each, off: off + r] = b.float().cpu() * scales[i]
        off += r
    return seg
