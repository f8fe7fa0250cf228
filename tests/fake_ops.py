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
              out_ld=None, max_ctas=0, workspace=None, cta_pair=None, gn_stat=None):
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
    if gn_stat is not None:
        # b200_conv_gemm_gnstat: per image / 32-pixel slab of a tile / 4-channel unit (sum, sum of squares) of the STORED values
        from audioldm_with_lora_b200.ops import box_rows
        assert stride == 1 and out.dtype == torch.bfloat16
        bh, _ = box_rows(h, w, nb)
        spi = w * bh // 32
        v = out.view(nb, h, w, -1)[..., :pw.n_valid].float().reshape(nb, h, w, pw.n_valid // 4, 4)
        hh = torch.arange(h).view(h, 1).expand(h, w)
        ww = torch.arange(w).view(1, w).expand(h, w)
        slab = ((hh // bh) * spi + ((hh % bh) * w + ww) // 32).reshape(-1)
        st = gn_stat.view(nb, -1, pw.n_valid // 4, 2)
        st.zero_()
        flat = v.reshape(nb, h * w, pw.n_valid // 4, 4)
        st[..., 0].index_add_(1, slab, flat.sum(-1))
        st[..., 1].index_add_(1, slab, (flat * flat).sum(-1))
    return out


def gn_stat_slabs(nb, h, w):
    from audioldm_with_lora_b200.ops import box_rows
    bh, _ = box_rows(h, w, nb)
    return ((h + bh - 1) // bh) * (w * bh // 32) if (w * bh) % 32 == 0 else 0


def groupnorm_apply(x0, c0, st0, x1, c1, st1, nb, hw, gamma, beta, eps, silu, y, groups=32):
    c = c0 + c1
    s = st0.view(nb, -1, c0 // 4, 2).double().sum(1)
    if c1:
        s = torch.cat([s, st1.view(nb, -1, c1 // 4, 2).double().sum(1)], 1)
    g = s.view(nb, groups, c // 4 // groups, 2).sum(2)
    cnt = hw * (c // groups)
    mean = g[..., 0] / cnt
    var = (g[..., 1] / cnt - mean * mean).clamp_min(0)
    x = torch.cat([x0.view(nb, hw, c0)] + ([x1.view(nb, hw, c1)] if c1 else []), -1).float()
    xn = (x.view(nb, hw, groups, -1) - mean.float()[:, None, :, None]) * torch.rsqrt(var.float() + eps)[:, None, :, None]
    out = xn.view(nb, hw, c) * gamma + beta
    if silu:
        out = F.silu(out)
    y.copy_(out.view(y.shape).to(y.dtype))
    return y


def set_sm_budget(n):
    pass


def linear_stats(pw, x, m, out, stat_out, *, a1=None, down=None, residual=None, max_ctas=0):
    if down is not None:
        linear_lora(pw, down, x, m, out, residual=residual)
    else:
        conv_gemm(pw, x, 1, m, 1, out, a1=a1, residual=residual)
    o = out.view(m, pw.n_valid // 64, 64).float()
    st = stat_out.view(m, pw.n_valid // 64, 2)
    st[..., 0] = o.sum(-1)
    st[..., 1] = (o * o).sum(-1)
    return out


def linear_ln(pw, x, m, out, stats, *, eps=1e-5, down=None, residual=None, max_ctas=0):
    """Emulates the kernel's algebra (raw x through gamma-folded weights, per-row correction in the epilogue, row
    statistics from the producer's per-chunk partial sums)."""
    c = pw.c0
    xf = x.view(m, c).float()
    st = stats.view(m, c // 64, 2).sum(1)
    mu = (st[:, 0] / c)[:, None]
    rs = ((st[:, 1] / c)[:, None] - mu * mu).clamp_min(0).add(eps).rsqrt()
    acc = xf @ pw.w[:, :c].float().T
    if down is not None:
        r = down.lora_rows
        t = xf @ down.w[:r].float().T - mu * down.ln_g[:r] + down.bias[:r] / rs
        T = torch.zeros(m, 64); T[:, :r] = t
        acc = acc + T.to(bf16).float() @ pw.w[:, c:c + 64].float().T
    acc = (acc - mu * pw.ln_g) * rs + pw.bias
    if pw.geglu:
        tt = acc.view(m, -1, 2, pw.block_n // 2)
        acc = (tt[:, :, 0] * F.gelu(tt[:, :, 1])).reshape(m, -1)
    acc = acc[:, :pw.n_valid]
    if residual is not None:
        acc = acc + residual.reshape(m, -1)[:, :pw.n_valid].float()
    _store(out, acc, pw.n_valid)
    return out


def linear_lora_ok(pw, down):
    return (getattr(pw, "lora_fused", False) and down is not None and getattr(down, "lora_rows", 0) > 0 and pw.c1 == 64 and
            pw.c2 == 0 and pw.ntaps == 1 and not pw.geglu and pw.ksplit == 1 and pw.block_n <= 192)


def linear_lora(pw, down, x, m, out, *, residual=None, t_out=None, max_ctas=0):
    xf = x.view(m, pw.c0).float()
    T = torch.zeros(m, 64)
    T[:, :down.lora_rows] = xf @ down.w[:down.lora_rows].float().T
    T = T.to(bf16)
    if t_out is not None:
        t_out.view(m, 64).copy_(T)
    conv_gemm(pw, x, 1, m, 1, out, a1=T, residual=residual)
    return out


def groupnorm_silu(x0, c0, x1, c1, nb, hw, gamma, beta, eps, silu, y, groups=32):
    x = x0.view(nb, hw, c0).float()
    if c1:
        x = torch.cat([x, x1.view(nb, hw, c1).float()], -1)
    r = F.group_norm(x.permute(0, 2, 1), groups, gamma, beta, eps)
    if silu:
        r = F.silu(r)
    y.reshape(-1)[: nb * hw * (c0 + c1)].view(nb, hw, c0 + c1).copy_(r.permute(0, 2, 1).to(bf16))    # (y may carry slack rows)
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


# ------------------------------------------------------------------------------------------ fine-tuning step
def _from_ptr(ptr, n, dtype):
    """A torch view of `n` elements of CPU memory at address `ptr` (the descriptor-table kernels take raw pointers)."""
    import ctypes
    import numpy as np
    if dtype == bf16:
        buf = (ctypes.c_uint16 * n).from_address(ptr)
        return torch.from_numpy(np.frombuffer(buf, dtype=np.uint16)).view(bf16)
    buf = (ctypes.c_float * n).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=np.float32))


def attention_lse(qkv, out, lse, batch, seq, heads, head_dim, scale=None):
    scale = head_dim ** -0.5 if scale is None else scale
    q, k, v = [t.view(batch, seq, heads, head_dim).transpose(1, 2).float()
               for t in qkv.view(batch, seq, -1).chunk(3, -1)]
    s = (q @ k.transpose(-1, -2)) * scale
    lse.view(batch, heads, seq).copy_(torch.logsumexp(s, -1) * 1.4426950408889634)
    o = torch.softmax(s, -1) @ v
    out.view(batch, seq, heads * head_dim).copy_(o.transpose(1, 2).reshape(batch, seq, -1).to(bf16))
    return out


def attention_bwd(qkv, o, dout, lse, delta, dqkv, batch, seq, heads, head_dim, scale=None):
    scale = head_dim ** -0.5 if scale is None else scale
    C = heads * head_dim
    q, k, v = [t.view(batch, seq, heads, head_dim).transpose(1, 2).float()
               for t in qkv.view(batch, seq, -1).chunk(3, -1)]
    do = dout.view(batch, seq, heads, head_dim).transpose(1, 2).float()
    of = o.view(batch, seq, heads, head_dim).transpose(1, 2).float()
    d = (do * of).sum(-1)
    delta.view(batch, heads, seq).copy_(d)
    p = torch.exp2((q @ k.transpose(-1, -2)) * scale * 1.4426950408889634 - lse.view(batch, heads, seq, 1))
    dv = p.transpose(-1, -2) @ do
    ds = p * (do @ v.transpose(-1, -2) - d[..., None])
    dq, dk = (ds @ k) * scale, (ds.transpose(-1, -2) @ q) * scale
    out = torch.cat([t.transpose(1, 2).reshape(batch, seq, C) for t in (dq, dk, dv)], -1)
    dqkv.view(batch, seq, 3 * C).copy_(out.to(bf16))
    return dqkv


def groupnorm_silu_stats(x0, c0, x1, c1, nb, hw, gamma, beta, eps, silu, y, stats, groups=32):
    groupnorm_silu(x0, c0, x1, c1, nb, hw, gamma, beta, eps, silu, y, groups)
    x = x0.view(nb, hw, c0).float()
    if c1:
        x = torch.cat([x, x1.view(nb, hw, c1).float()], -1)
    xg = x.view(nb, hw, groups, -1).permute(0, 2, 1, 3).reshape(nb, groups, -1)
    st = stats.view(nb, groups, 2)
    st[..., 0] = xg.mean(-1)
    st[..., 1] = (xg.var(-1, unbiased=False) + eps).rsqrt()
    return y


def groupnorm_silu_bwd(x0, c0, x1, c1, nb, hw, gamma, beta, stats, silu, dy, dres, res_ld, dx0, dx1, groups=32):
    """Closed form from the SAVED statistics (what the kernel does), not autograd of F.group_norm."""
    C = c0 + c1
    cpg = C // groups
    x = x0.view(nb, hw, c0).float()
    if c1:
        x = torch.cat([x, x1.view(nb, hw, c1).float()], -1)
    st = stats.view(nb, groups, 2)
    mean = st[..., 0].repeat_interleave(cpg, 1)[:, None, :]
    rstd = st[..., 1].repeat_interleave(cpg, 1)[:, None, :]
    xh = (x - mean) * rstd
    g = dy.view(nb, hw, C).float()
    if silu:
        yh = xh * gamma + beta
        sg = torch.sigmoid(yh)
        g = g * sg * (1 + yh * (1 - sg))
    dxh = g * gamma
    m1 = dxh.view(nb, hw, groups, cpg).mean((1, 3)).repeat_interleave(cpg, 1)[:, None, :]
    m2 = (dxh * xh).view(nb, hw, groups, cpg).mean((1, 3)).repeat_interleave(cpg, 1)[:, None, :]
    gx = rstd * (dxh - m1 - xh * m2)
    if dres is not None:
        gx = gx + dres.view(nb, hw, res_ld)[:, :, :C].float()
    dx0.view(nb, hw, c0).copy_(gx[..., :c0].to(bf16))
    if dx1 is not None:
        dx1.view(nb, hw, c1).copy_(gx[..., c0:].to(bf16))


@torch.enable_grad()          # may run inside an autograd.Function.backward (grad mode off)
def layernorm_bwd(x, dy, m, c, gamma, eps, dres, dx):
    xr = x.view(m, c).float().requires_grad_(True)
    F.layer_norm(xr, (c,), gamma, torch.zeros_like(gamma), eps).backward(dy.view(m, c).float())
    g = xr.grad + (dres.view(m, c).float() if dres is not None else 0.0)
    dx.view(m, c).copy_(g.to(bf16))
    return dx


def geglu_fwd(h, m, f, out):
    val, gate = h.view(m, 2 * f).float().chunk(2, -1)
    out.view(m, f).copy_((val * F.gelu(gate)).to(bf16))
    return out


@torch.enable_grad()          # may run inside an autograd.Function.backward (grad mode off)
def geglu_bwd(h, dout, m, f, dh):
    hr = h.view(m, 2 * f).float().requires_grad_(True)
    val, gate = hr.chunk(2, -1)
    (val * F.gelu(gate)).backward(dout.view(m, f).float())
    dh.view(m, 2 * f).copy_(hr.grad.to(bf16))
    return dh


def lora_wgrad(descs, m):
    for d in descs:
        u = _from_ptr(d.u, (m - 1) * d.ldu + d.C, bf16).float()
        v = _from_ptr(d.v, (m - 1) * d.ldv + d.r, bf16).float()
        U = torch.as_strided(u, (m, d.C), (d.ldu, 1))
        V = torch.as_strided(v, (m, d.r), (d.ldv, 1))
        G = (U.T @ V) * d.scale                                   # [C, r]
        span = (d.C - 1) * d.ldc + (d.r - 1) * d.ldj + 1
        out = torch.as_strided(_from_ptr(d.out, span, torch.float32), (d.C, d.r), (d.ldc, d.ldj))
        out += G


def zero_insert(dy, nb, h, w, c, z):
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    zz = z.view(nb, h, w, c)
    zz.zero_()
    zz[:, ::2, ::2] = dy.view(nb, ho, wo, c)
    return z


@torch.enable_grad()          # may run inside an autograd.Function.backward (grad mode off)
def upsample_nearest_bwd(dy, nb, h, w, c, ho, wo, dx):
    xr = torch.zeros(nb, c, h, w, requires_grad=True)
    F.interpolate(xr, size=(ho, wo), mode="nearest").backward(dy.view(nb, ho, wo, c).float().permute(0, 3, 1, 2))
    dx.view(nb, h, w, c).copy_(xr.grad.permute(0, 2, 3, 1).to(bf16))
    return dx


def add_bf16(y, x):
    y.copy_((y.float() + x.float()).to(bf16))
    return y


def mse_grad(pred_nhwc, noise_nchw, nb, hw, c_pad, inv_count, loss_sum, deps):
    diff = pred_nhwc.view(nb, hw, 8) - noise_nchw.view(nb, 8, hw).permute(0, 2, 1)
    loss_sum += (diff ** 2).sum()
    d = deps.view(nb, hw, c_pad)
    d.zero_()
    d[..., :8] = (2 * diff * inv_count).to(bf16)


def lora_refresh(descs_dev, n, flat):
    import numpy as np
    from audioldm_with_lora_b200.ops import REFRESH_DTYPE
    recs = np.frombuffer(descs_dev.numpy().tobytes(), dtype=REFRESH_DTYPE)
    for r in recs[:n]:
        sr, sc, ld = int(r["src_rows"]), int(r["src_cols"]), int(r["dst_ld"])
        src = flat[int(r["src_off"]): int(r["src_off"]) + sr * sc].view(sr, sc) * float(r["scale"])
        if r["transpose"]:
            src = src.t()
        rows, cols = src.shape
        dst = torch.as_strided(_from_ptr(int(r["dst"]), (rows - 1) * ld + cols, bf16), (rows, cols), (ld, 1))
        dst.copy_(src.to(bf16))


TRAIN_OPS = ("attention_lse", "attention_bwd", "groupnorm_silu_stats", "groupnorm_silu_bwd", "layernorm_bwd",
             "geglu_fwd", "geglu_bwd", "lora_wgrad", "zero_insert", "upsample_nearest_bwd", "add_bf16", "mse_grad",
             "lora_refresh")


def install(monkeypatch, ops_module):
    """Replace the kernel wrappers of `ops_module` (keeps PackedWeight / tiling helpers)."""
    for name in ("conv_gemm", "groupnorm_silu", "layernorm", "attention", "time_class_embed",
                 "pack_nchw_to_nhwc", "unpack_nhwc_to_nchw", "upsample_nearest", "sampler_step", "add_noise",
                 "adamw_flat", "mse_partial", "set_sm_budget", "gn_stat_slabs", "groupnorm_apply", "linear_lora_ok", "linear_lora", "linear_ln", "linear_stats",
                 "conv1d", "lrelu_mean3", "embed_layernorm", "f32_to_bf16", "gemm_nt", "softmax_rows", "conv3x3_s2_pad01") + TRAIN_OPS:
        monkeypatch.setattr(ops_module, name, globals()[name])


# ------------------------------------------------------------------------------------------ HiFi-GAN vocoder
def conv1d(pw, x, nb, length, out, *, dh0, dh_step, m_rows=None, residual=None, res_slope=1.0, act_slope=1.0,
           act_tanh=False, out_ld=None, out_batch_stride=0, cta_pair=None):
    """b200_conv1d semantics: out[n, q] = act(bias + sum_t x[n, q + dh0 + t dh_step] W_t^T + unact(residual[n, q]))."""
    rows = length if m_rows is None else m_rows
    c = pw.c0
    xf = x.view(nb, length, c).float()
    acc = torch.zeros(nb, rows, pw.n_pad)
    wk = pw.w.float().view(pw.n_pad, pw.ntaps, c)
    q = torch.arange(rows)
    for t in range(pw.ntaps):
        src = q + dh0 + t * dh_step
        ok = (src >= 0) & (src < length)
        g = torch.zeros(nb, rows, c)
        g[:, ok] = xf[:, src[ok]]
        acc += g @ wk[:, t].T
    if pw.bias is not None:
        acc = acc + pw.bias
    acc = acc[..., :pw.n_valid]
    if residual is not None:
        r = residual.view(nb, length, pw.n_valid).float()
        acc = acc + torch.minimum(r, r / res_slope)
    if int(act_tanh) == 2:
        acc = torch.nn.functional.gelu(acc)
    else:
        acc = torch.tanh(acc) if act_tanh else torch.maximum(acc, acc * act_slope)
    ld = out_ld if out_ld is not None else pw.n_valid
    bs = out_batch_stride if out_batch_stride else rows * ld
    flat = out.view(-1)
    idx = (torch.arange(nb)[:, None, None] * bs + torch.arange(rows)[None, :, None] * ld + torch.arange(pw.n_valid)[None, None, :])
    flat[idx.reshape(-1)] = acc.reshape(-1).to(out.dtype)
    return out


def embed_layernorm(ids, pos_ids, word, pos, type0, gamma, beta, eps, y):
    e = (word[ids.long()] + type0) + pos[pos_ids.long()]
    y.copy_(torch.nn.functional.layer_norm(e, (word.shape[1],), gamma, beta, eps).to(y.dtype).view_as(y))
    return y


def lrelu_mean3(a0, a1, a2, in_slope, out_slope, y):
    xs = [torch.minimum(a.float(), a.float() / in_slope) for a in (a0, a1, a2)]
    m = (xs[0] + xs[1] + xs[2]) / 3.0
    y.copy_(torch.maximum(m, m * out_slope).to(y.dtype))
    return y


def f32_to_bf16(x, y):
    y.copy_(x.to(y.dtype))
    return y


# ------------------------------------------------------------------------------------------ VAE
def gemm_nt(a, b, n_valid, out, block_n, out_ld=None):
    acc = a.float() @ b[:n_valid].float().T
    ld = out_ld if out_ld is not None else n_valid
    o = torch.as_strided(out, (a.shape[0], n_valid), (ld, 1), out.storage_offset())
    o.copy_(acc.to(out.dtype))
    return out


def softmax_rows(s, rows, cols, cols_pad, p, scale=1.0):
    pr = F.softmax(s[:rows, :cols].float() * scale, dim=-1)
    p[:rows, :cols_pad] = 0
    p[:rows, :cols] = pr.to(p.dtype)
    return p


def conv3x3_s2_pad01(pw, x, nb, h, w, out, cta_pair=None):
    """b200_conv3x3_s2_pad01 semantics: F.pad(x, (0, 1, 0, 1)) + conv k3 s2 p0 over NHWC bf16."""
    c = pw.c0
    xn = x.view(nb, h, w, c).float().permute(0, 3, 1, 2)
    wt = pw.w.float().view(pw.n_pad, 3, 3, c).permute(0, 3, 1, 2)[:pw.n_valid]
    y = F.conv2d(F.pad(xn, (0, 1, 0, 1)), wt, pw.bias[:pw.n_valid] if pw.bias is not None else None, stride=2)
    out.view(nb, y.shape[2], y.shape[3], pw.n_valid).copy_(y.permute(0, 2, 3, 1).to(out.dtype))
    return out
