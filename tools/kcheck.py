#!/usr/bin/env python
"""One-shot GPU kernel checks, one case per process (a CUDA fault cannot poison the next case).

    python tools/kcheck.py <case> [k=v ...]      -> one JSON line on stdout

References here are torch fp32 ops on bf16-rounded inputs (per-kernel numerics); the oracle-level
parity lives in tests/.  Used for bring-up on the GPU box; results land in gpurun_out/kcheck.jsonl.
"""
from __future__ import annotations

import json
import math
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops, packing  # noqa: E402

DEV = "cuda"
bf16 = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def case_linear(m=1000, n=256, k=256, bn=0, bias=1, resid=1, fp32=0):
    a = rnd(m, k, seed=1).to(bf16)
    w = rnd(n, k, seed=2, scale=k ** -0.5)
    b = rnd(n, seed=3) if bias else None
    r = rnd(m, n, seed=4).to(bf16) if resid else None
    bn = bn or ops.choose_block_n(n, math.ceil(m / 128))
    pw = packing.pack([w.cpu()], b.cpu() if bias else None, bn, 1, k, device=DEV)
    out = torch.full((m, n), float("nan"), dtype=torch.float32 if fp32 else bf16, device=DEV)
    ops.conv_gemm(pw, a, 1, m, 1, out, residual=r)
    torch.cuda.synchronize()
    ref = a.float() @ w.to(bf16).float().T
    if bias:
        ref = ref + b
    if resid:
        ref = ref + r.float()
    return {"rel": rel(out, ref), "bn": bn, "nan": int(torch.isnan(out.float()).sum())}


def case_conv(nb=2, h=20, w=16, ci=128, co=128, stride=1, bn=0, rowvec=1, resid=0, seed=0):
    x = rnd(nb, ci, h, w, seed=seed + 1).to(bf16)
    wt = rnd(co, ci, 3, 3, seed=seed + 2, scale=(9 * ci) ** -0.5)
    b = rnd(co, seed=seed + 3)
    rv = rnd(nb, co, seed=seed + 4) if rowvec else None
    ho, wo = ((h - 1) // 2 + 1, (w - 1) // 2 + 1) if stride == 2 else (h, w)
    r = rnd(nb, co, ho, wo, seed=seed + 5).to(bf16) if resid else None
    bn = bn or ops.choose_block_n(co, ops.num_m_tiles(nb, h, w))
    wk = packing.conv3x3_to_k(wt.cpu())
    pw = packing.pack([wk], b.cpu(), bn, 9, wk.shape[1] // 9, device=DEV)
    xin = x.permute(0, 2, 3, 1).contiguous()
    if pw.c0 != ci:
        xp = torch.zeros(nb, h, w, pw.c0, dtype=bf16, device=DEV)
        xp[..., :ci] = xin
        xin = xp
    out = torch.full((nb * ho * wo, co), float("nan"), dtype=bf16, device=DEV)
    rr = r.permute(0, 2, 3, 1).contiguous() if resid else None
    ops.conv_gemm(pw, xin, nb, h, w, out, stride=stride, rowvec=rv, rowvec_ld=co, residual=rr)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), wt.to(bf16).float(), b, stride=stride, padding=1)
    if rowvec:
        ref = ref + rv[:, :, None, None]
    if resid:
        ref = ref + r.float()
    got = out.view(nb, ho, wo, co).permute(0, 3, 1, 2)
    return {"rel": rel(got, ref), "bn": bn, "nan": int(torch.isnan(out.float()).sum())}


def case_segments(nb=2, h=20, w=16, c=128, ch=128, cs=64, co=128):
    """conv2 (3x3 over n2) + 1x1 shortcut over cat([xh, xs]) in one accumulator."""
    n2 = rnd(nb, c, h, w, seed=1).to(bf16)
    xh = rnd(nb, ch, h, w, seed=2).to(bf16)
    xs = rnd(nb, cs, h, w, seed=3).to(bf16)
    w2 = rnd(co, c, 3, 3, seed=4, scale=(9 * c) ** -0.5)
    ws = rnd(co, ch + cs, 1, 1, seed=5, scale=(ch + cs) ** -0.5)
    b = rnd(co, seed=6)
    bn = ops.choose_block_n(co, ops.num_m_tiles(nb, h, w))
    pw = packing.pack([packing.conv3x3_to_k(w2.cpu()), ws[:, :ch, 0, 0].cpu(), ws[:, ch:, 0, 0].cpu()], b.cpu(), bn, 9, c,
                      ch, cs, device=DEV)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    out = torch.full((nb * h * w, co), float("nan"), dtype=bf16, device=DEV)
    ops.conv_gemm(pw, nhwc(n2), nb, h, w, out, a1=nhwc(xh), a2=nhwc(xs))
    torch.cuda.synchronize()
    ref = F.conv2d(n2.float(), w2.to(bf16).float(), b, padding=1) + F.conv2d(torch.cat([xh, xs], 1).float(),
                                                                             ws.to(bf16).float())
    return {"rel": rel(out.view(nb, h, w, co).permute(0, 3, 1, 2), ref), "bn": bn}


def case_geglu(m=1000, c=256, bn=0):
    a = rnd(m, c, seed=1).to(bf16)
    w = rnd(8 * c, c, seed=2, scale=c ** -0.5)
    b = rnd(8 * c, seed=3)
    bn = bn or ops.choose_block_n(8 * c, math.ceil(m / 128), geglu=True)
    pw = packing.pack([w.cpu()], b.cpu(), bn, 1, c, geglu=True, device=DEV)
    out = torch.full((m, 4 * c), float("nan"), dtype=bf16, device=DEV)
    ops.conv_gemm(pw, a, 1, m, 1, out)
    torch.cuda.synchronize()
    hcat = a.float() @ w.to(bf16).float().T + b
    val, gate = hcat.chunk(2, -1)
    return {"rel": rel(out, val * F.gelu(gate)), "bn": bn, "nan": int(torch.isnan(out.float()).sum())}


def case_gn(nb=2, hw=320, c0=128, c1=0, silu=1, eps=1e-5):
    x0 = rnd(nb, hw, c0, seed=1).to(bf16) + 0.5
    x1 = rnd(nb, hw, c1, seed=2).to(bf16) * 2 if c1 else None
    c = c0 + c1
    gamma, beta = rnd(c, seed=3) * 0.1 + 1, rnd(c, seed=4) * 0.1
    y = torch.empty(nb, hw, c, dtype=bf16, device=DEV)
    ops.groupnorm_silu(x0, c0, x1, c1, nb, hw, gamma, beta, eps, bool(silu), y)
    torch.cuda.synchronize()
    xc = torch.cat([x0, x1], -1) if c1 else x0
    ref = F.group_norm(xc.float().permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    return {"rel": rel(y.permute(0, 2, 1), ref)}


def case_ln(m=1000, c=384):
    x = rnd(m, c, seed=1).to(bf16) + 0.3
    gamma, beta = rnd(c, seed=3) * 0.1 + 1, rnd(c, seed=4) * 0.1
    y = torch.empty(m, c, dtype=bf16, device=DEV)
    ops.layernorm(x, m, c, gamma, beta, 1e-5, y)
    torch.cuda.synchronize()
    return {"rel": rel(y, F.layer_norm(x.float(), (c,), gamma, beta, 1e-5))}


def case_attn(b=2, s=300, heads=8, d=32, variant=0, time=0):
    qkv = rnd(b, s, 3 * heads * d, seed=1).to(bf16)
    out = torch.full((b, s, heads * d), float("nan"), dtype=bf16, device=DEV)
    ops.attention(qkv, out, b, s, heads, d, variant=variant)
    torch.cuda.synchronize()
    q, k, v = [t.view(b, s, heads, d).transpose(1, 2).float() for t in qkv.chunk(3, -1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s, heads * d)
    res = {"rel": rel(out, ref), "nan": int(torch.isnan(out.float()).sum()), "variant": variant}
    if time:
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.attention(qkv, out, b, s, heads, d, variant=variant)
        t0.record()
        for _ in range(10):
            ops.attention(qkv, out, b, s, heads, d, variant=variant)
        t1.record(); torch.cuda.synchronize()
        res["ms"] = t0.elapsed_time(t1) / 10
    return res


def case_upsample(nb=2, h=32, w=2, c=64, ho=63, wo=4):
    x = rnd(nb, h, w, c, seed=1).to(bf16)
    y = torch.empty(nb, ho, wo, c, dtype=bf16, device=DEV)
    ops.upsample_nearest(x, nb, h, w, c, ho, wo, y)
    torch.cuda.synchronize()
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), size=(ho, wo), mode="nearest").permute(0, 2, 3, 1)
    return {"maxabs": (y.float() - ref).abs().max().item()}


def case_unet(arch="S", nb=2, h=32, r=8, lora=1, variant=0, time=0):
    """Full UNet forward vs the fp32 CPU oracle, with per-layer taps."""
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.arch import CONFIGS
    from audioldm_with_lora_b200.engine import UNetEngine
    from audioldm_with_lora_b200.lora import parse_lora_state_dict
    from oracle import unet_ref
    cfg = CONFIGS[arch]
    sd = synthetic.random_unet_state_dict(cfg, seed=0)
    eng = UNetEngine(cfg, sd, DEV)
    eng.attn_variant = variant
    ora_lora = None
    if lora:
        ad = parse_lora_state_dict(synthetic.random_lora_state_dict(cfg, r, fmt="peft"))
        eng.set_lora(ad, 1.0)
        ora_lora = unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()})
    x = synthetic.initial_latents(nb, h)
    pos, neg = synthetic.clap_embeddings(nb)
    t = 501
    taps_o, taps_e = {}, {}
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, unet_ref.ARCHS[arch], x, t, pos, lora=ora_lora, taps=taps_o)
    out = eng.forward(x, t, pos, taps=taps_e)
    torch.cuda.synchronize()
    res = {"rel": rel(out.cpu(), ref), "nan": int(torch.isnan(out).sum())}
    tl = {}
    for k, v in taps_o.items():
        if k in taps_e:
            tl[k] = round(rel(taps_e[k].cpu(), v), 5)
    res["taps"] = tl
    if time:
        for _ in range(2):
            eng.forward(x, t, pos)
        torch.cuda.synchronize()
        t0 = __import__("time").time()
        for _ in range(5):
            eng.forward(x, t, pos)
        torch.cuda.synchronize()
        res["ms_eager"] = (__import__("time").time() - t0) / 5 * 1e3
    return res


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}


def main():
    name = sys.argv[1]
    kw = {}
    for a in sys.argv[2:]:
        k, v = a.split("=")
        kw[k] = v if v.isalpha() else (float(v) if "." in v or "e" in v else int(v))
    t0 = time.time()
    rec = {"case": name, "args": kw}
    try:
        rec.update(CASES[name](**kw))
        rec["ok"] = True
    except Exception as e:  # noqa: BLE001
        rec["ok"] = False
        rec["error"] = f"{type(e).__name__}: {e}"[:600]
    rec["sec"] = round(time.time() - t0, 2)
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
