"""TEST INFRASTRUCTURE: torch-CPU emulation of the C-ABI kernel *semantics* (include/b200ldm.h).

Lets the CPU suite exercise the host logic (weight packing, K-segment ordering, arena, graph
scheduling, LoRA folding) against the oracle without a GPU.  It is never imported by the package;
tests monkeypatch `audioldm_with_lora_b200.ops` with it.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

bf16 = torch.bfloat16


def _store(out, val, n_valid, ld=None):
    o2 = out.view(-1, ld if ld else out.shape[-1])
    o2[:, :n_valid] = val.to(out.dtype)


def conv_gemm(pw, a0, nb, h, w, out, *, a1=None, a2=None, stride=1, rowvec=None, rowvec_ld=0, residual=None,
              out_ld=None, max_ctas=0, workspace=None, cta_pair=None):
    M = nb * h * w
    x = a0.view(nb, h, w, pw.c0).float()
    cols = []
    if pw.ntaps == 9:
        xp = F.pad(x, (0, 0, 1, 1, 1, 1))
        for kh in range(3):
            for kw in range(3):
                cols.append(xp[:, kh:kh + h, kw:kw + w, :].reshape(M, pw.c0))
    else:
        cols.append(x.reshape(M, pw.c0))
    if pw.c1:
        cols.append(a1.view(M, pw.c1).float())
    if pw.c2:
        cols.append(a2.view(M, pw.c2).float())
    A = torch.cat(cols, 1)
    assert A.shape[1] == pw.k, (A.shape, pw.k)
    acc = A @ pw.w.float().T
    if pw.bias is not None:
        acc = acc + pw.bias
    if pw.geglu:
        t = acc.view(M, -1, 2, pw.block_n // 2)
        acc = (t[:, :, 0] * F.gelu(t[:, :, 1])).reshape(M, -1)
    acc = acc[:, :pw.n_valid]
    if rowvec is not None:
        acc = (acc.view(nb, h * w, -1) + rowvec[:, None, :pw.n_valid]).view(M, -1)
    if stride == 2:
        acc = acc.view(nb, h, w, -1)[:, ::2, ::2].reshape(-1, pw.n_valid)
    if residual is not None:
        acc = acc + residual.reshape(acc.shape[0], -1)[:, :pw.n_valid].float()
    _store(out, acc, pw.n_valid, out_ld)
    return out


def groupnorm_silu(x0, c0, x1, c1, nb, hw, gamma, beta, eps, silu, y, groups=32):
    x = x0.view(nb, hw, c0).float()
    if c1:
        x = torch.cat([x, x1.view(nb, hw, c1).float()], -1)
    r = F.group_norm(x.permute(0, 2, 1), groups, gamma, beta, eps)
    if silu:
        r = F.silu(r)
    y.view(nb, hw, c0 + c1).copy_(r.permute(0, 2, 1).to(bf16))
    return y


def layernorm(x, m, c, gamma, beta, eps, y):
    y.view(m, c).copy_(F.layer_norm(x.view(m, c).float(), (c,), gamma, beta, eps).to(bf16))
    return y


def attention(qkv, out, batch, seq, heads, head_dim, scale=None, variant=0):
    q, k, v = [t.view(batch, seq, heads, head_dim).transpose(1, 2).float()
               for t in qkv.view(batch, seq, -1).chunk(3, -1)]
    o = F.scaled_dot_product_attention(q, k, v, scale=scale)
    out.view(batch, seq, heads * head_dim).copy_(o.transpose(1, 2).reshape(batch, seq, -1).to(bf16))
    return out


def time_class_embed(t_steps, step_ptr, per_sample, labels, nb, tproj, ted, class_in, w1, b1, w2, b2, wc, bc, emb,
                     silu_emb):
    t = t_steps[:nb] if per_sample else t_steps[int(step_ptr.item()) if step_ptr is not None else 0].expand(nb)
    half = tproj // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = t[:, None].float() * freqs[None]
    e = torch.cat([torch.cos(arg), torch.sin(arg)], -1)
    e = F.linear(F.silu(F.linear(e, w1.t(), b1)), w2.t(), b2)       # weights arrive transposed ([in, out])
    c = F.linear(labels, wc.t(), bc)
    full = torch.cat([e, c], -1)
    if emb is not None:
        emb.copy_(full)
    silu_emb.copy_(F.silu(full).to(bf16))


def pack_nchw_to_nhwc(x, nb, c, hw, c_pad, y):
    y.view(nb, hw, c_pad)[:, :, :c] = x.view(nb, c, hw).permute(0, 2, 1).to(bf16)
    return y


def unpack_nhwc_to_nchw(x, nb, c, hw, y):
    y.view(nb, c, hw).copy_(x.view(nb, hw, c).permute(0, 2, 1))
    return y


def upsample_nearest(x, nb, h, w, c, ho, wo, y):
    r = F.interpolate(x.view(nb, h, w, c).permute(0, 3, 1, 2).float(), size=(ho, wo), mode="nearest")
    y.view(nb, ho, wo, c).copy_(r.permute(0, 2, 3, 1).to(bf16))
    return y


def sampler_step(eps, x, x_saved, hist, table, step_ptr, guidance, do_cfg, nb, hw, c, c_pad, xin_next):
    step = int(step_ptr.item())
    row = table.view(-1, 8)[step]
    a, b = row[0].item(), row[1].item()
    wts = [row[2 + i].item() for i in range(4)]
    flags = int(row[6:7].view(torch.int32).item())
    slots = int(row[7:8].view(torch.int32).item())
    n = nb * hw * c
    e_all = eps.view(-1)
    e = e_all[:n] + guidance * (e_all[n:2 * n] - e_all[:n]) if do_cfg else e_all[:n].clone()
    eh = wts[0] * e
    for i, sh in enumerate((0, 4, 8)):
        if wts[i + 1] != 0.0:
            eh = eh + wts[i + 1] * hist.view(-1, n)[(slots >> sh) & 15]
    push = ((flags >> 4) & 7) - 1
    if push >= 0:
        hist.view(-1, n)[push] = e
    xf = x.view(-1)
    if flags & 2:
        x_saved.view(-1).copy_(xf)
    xb = x_saved.view(-1).clone() if flags & 1 else xf.clone()
    xn = a * xb + b * eh
    xf.copy_(xn)
    if xin_next is not None:
        v = xn.view(nb, hw, c).to(bf16)
        xv = xin_next.view(-1, hw, c_pad)
        xv[:nb, :, :c] = v
        if do_cfg:
            xv[nb:2 * nb, :, :c] = v
    step_ptr += 1


def add_noise(x0, noise, sqrt_ac, sqrt_1mac, out):
    shp = (-1,) + (1,) * (x0.dim() - 1)
    out.copy_(sqrt_ac.view(shp) * x0 + sqrt_1mac.view(shp) * noise)
    return out


def adamw_flat(param, grad, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    g = grad * grad_scale
    param.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    param.addcdiv_(m, denom, value=-lr / bc1)


def mse_partial(pred, target, out_sum):
    out_sum += ((pred - target) ** 2).sum()


def install(monkeypatch, ops_module):
    """Replace the kernel wrappers of `ops_module` (keeps PackedWeight / tiling helpers)."""
    for name in ("conv_gemm", "groupnorm_silu", "layernorm", "attention", "time_class_embed",
                 "pack_nchw_to_nhwc", "unpack_nhwc_to_nchw", "upsample_nearest", "sampler_step", "add_noise",
                 "adamw_flat", "mse_partial"):
        monkeypatch.setattr(ops_module, name, globals()[name])
