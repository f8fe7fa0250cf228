"""GPU parity tests (pytest -m gpu) of the fine-tuning step's backward kernels, through the C-ABI, against torch
autograd (fp32) of the same op on the same bf16-rounded inputs.

Tolerance: outputs are bf16 (2^-9 relative per element) of fp32-accumulated sums -> rel-L2 <= 5e-3 per kernel; the
attention backward recomputes P from bf16 Q/K and a bf16-P forward, so it is held to 1.5e-2.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16
DEV = "cuda"
KERNEL_TOL = 5e-3
ATTN_BWD_TOL = 1.5e-2


@pytest.fixture(scope="module", autouse=True)
def _need_cuda_and_lib():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the hot path has no CPU fallback")
    from audioldm_with_lora_b200 import _lib
    _lib.load()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("nb,hw,c0,c1,silu,eps,use_res,need_dx1", [
    (2, 320, 128, 0, True, 1e-5, True, True),
    (3, 252, 384, 256, True, 1e-5, True, True),        # two-source (concat), odd pixel count
    (4, 64, 640, 640, True, 1e-5, False, False),       # skip gradient not needed
    (2, 1000, 256, 0, False, 1e-6, True, True),        # Transformer2DModel.norm: no SiLU
    (32, 64, 640, 0, True, 1e-5, True, True),          # more images than co-resident clusters of 8
])
def test_groupnorm_silu_bwd(nb, hw, c0, c1, silu, eps, use_res, need_dx1):
    from audioldm_with_lora_b200 import ops
    C = c0 + c1
    x0 = rnd(nb, hw, c0, seed=1).to(bf16)
    x1 = rnd(nb, hw, c1, seed=2).to(bf16) if c1 else None
    gamma, beta = 1.0 + 0.2 * rnd(C, seed=3), 0.3 * rnd(C, seed=4)
    dy = rnd(nb, hw, C, seed=5).to(bf16)
    dres = rnd(nb, hw, C, seed=6).to(bf16) if use_res else None
    y = torch.empty(nb, hw, C, dtype=bf16, device=DEV)
    stats = torch.empty(nb, 32, 2, dtype=torch.float32, device=DEV)
    ops.groupnorm_silu_stats(x0, c0, x1, c1, nb, hw, gamma, beta, eps, silu, y, stats)
    dx0 = torch.full((nb, hw, c0), float("nan"), dtype=bf16, device=DEV)
    dx1 = torch.full((nb, hw, c1), float("nan"), dtype=bf16, device=DEV) if (c1 and need_dx1) else None
    ops.groupnorm_silu_bwd(x0, c0, x1, c1, nb, hw, gamma, beta, stats, silu, dy, dres, C, dx0, dx1)
    # torch reference: NCHW group_norm on the concatenation
    xc = torch.cat([x0, x1], -1) if c1 else x0
    xr = xc.float().permute(0, 2, 1).reshape(nb, C, hw, 1).requires_grad_(True)
    yr = F.group_norm(xr, 32, gamma, beta, eps)
    if silu:
        yr = F.silu(yr)
    # forward output and saved statistics
    assert rel(y, yr.detach().reshape(nb, C, hw).permute(0, 2, 1)) < KERNEL_TOL
    xg = xc.float().reshape(nb, hw, 32, C // 32).permute(0, 2, 1, 3).reshape(nb, 32, -1)
    assert rel(stats[..., 0], xg.mean(-1)) < 1e-3 or (stats[..., 0] - xg.mean(-1)).abs().max() < 1e-3
    assert rel(stats[..., 1], (xg.var(-1, unbiased=False) + eps).rsqrt()) < 1e-3
    yr.backward(dy.float().permute(0, 2, 1).reshape(nb, C, hw, 1))
    ref = xr.grad.reshape(nb, C, hw).permute(0, 2, 1)
    if use_res:
        ref = ref + dres.float()
    assert not torch.isnan(dx0.float()).any()
    assert rel(dx0, ref[..., :c0]) < KERNEL_TOL
    if dx1 is not None:
        assert rel(dx1, ref[..., c0:]) < KERNEL_TOL


@pytest.mark.parametrize("m,c,use_res", [(1000, 256, True), (252, 384, False), (64, 640, True), (37, 1280, True)])
def test_layernorm_bwd(m, c, use_res):
    from audioldm_with_lora_b200 import ops
    x = rnd(m, c, seed=1).to(bf16)
    dy = rnd(m, c, seed=2).to(bf16)
    gamma, beta = 1.0 + 0.2 * rnd(c, seed=3), 0.1 * rnd(c, seed=4)
    dres = rnd(m, c, seed=5).to(bf16) if use_res else None
    dx = torch.full((m, c), float("nan"), dtype=bf16, device=DEV)
    ops.layernorm_bwd(x, dy, m, c, gamma, 1e-5, dres, dx)
    xr = x.float().requires_grad_(True)
    F.layer_norm(xr, (c,), gamma, beta, 1e-5).backward(dy.float())
    ref = xr.grad + (dres.float() if use_res else 0.0)
    assert not torch.isnan(dx.float()).any()
    assert rel(dx, ref) < KERNEL_TOL


@pytest.mark.parametrize("m,f", [(1000, 1024), (252, 1536), (64, 2560)])
def test_geglu_fwd_bwd(m, f):
    from audioldm_with_lora_b200 import ops
    h = rnd(m, 2 * f, seed=1).to(bf16)
    dout = rnd(m, f, seed=2).to(bf16)
    out = torch.empty(m, f, dtype=bf16, device=DEV)
    dh = torch.empty(m, 2 * f, dtype=bf16, device=DEV)
    ops.geglu_fwd(h, m, f, out)
    ops.geglu_bwd(h, dout, m, f, dh)
    hr = h.float().requires_grad_(True)
    val, gate = hr.chunk(2, -1)
    o = val * F.gelu(gate)
    o.backward(dout.float())
    assert rel(out, o.detach()) < KERNEL_TOL
    assert rel(dh, hr.grad) < KERNEL_TOL


@pytest.mark.parametrize("m,c,r", [(1000, 256, 8), (4032, 384, 16), (64, 640, 2), (1024, 640, 32), (333, 1280, 8)])
def test_lora_wgrad(m, c, r):
    from audioldm_with_lora_b200 import ops
    from audioldm_with_lora_b200._lib import ptr
    u = rnd(m, 3 * c, seed=1).to(bf16)                      # e.g. d_qkv: the k slice is columns [c, 2c)
    v = rnd(m, 64, seed=2).to(bf16)                         # e.g. T_qkv: this adapter's columns [r, 2r)
    out_b = torch.zeros(c, r, dtype=torch.float32, device=DEV)      # dB layout [Cout, r]
    out_a = torch.zeros(r, c, dtype=torch.float32, device=DEV)      # dA layout [r, Cin]
    d1 = ops.WgradDesc(ptr(u) + 2 * c, ptr(v) + 2 * r, ptr(out_b), 3 * c, 64, c, r, r, 1, 0.5)
    d2 = ops.WgradDesc(ptr(u) + 2 * c, ptr(v) + 2 * r, ptr(out_a), 3 * c, 64, c, r, 1, c, 1.0)
    ops.lora_wgrad([d1, d2], m)
    ref = u[:, c:2 * c].float().T @ v[:, r:2 * r].float()
    assert rel(out_b, 0.5 * ref) < 1e-4
    assert rel(out_a, ref.T) < 1e-4
    ops.lora_wgrad([d1], m)                                 # accumulates
    assert rel(out_b, ref) < 1e-4


def test_zero_insert_and_upsample_bwd_and_add():
    from audioldm_with_lora_b200 import ops
    nb, h, w, c = 2, 63, 4, 64
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    dy = rnd(nb, ho, wo, c, seed=1).to(bf16)
    z = torch.full((nb, h, w, c), float("nan"), dtype=bf16, device=DEV)
    ops.zero_insert(dy, nb, h, w, c, z)
    ref = torch.zeros(nb, h, w, c, device=DEV)
    ref[:, ::2, ::2] = dy.float()
    assert torch.equal(z.float(), ref)
    # nearest upsample backward == autograd of F.interpolate(size=...)
    for (hi, wi, hq, wq) in [(32, 2, 63, 4), (63, 4, 125, 8), (125, 8, 250, 16), (32, 2, 64, 4)]:
        g = rnd(nb, hq, wq, c, seed=2).to(bf16)
        dx = torch.empty(nb, hi, wi, c, dtype=bf16, device=DEV)
        ops.upsample_nearest_bwd(g, nb, hi, wi, c, hq, wq, dx)
        xr = torch.zeros(nb, c, hi, wi, device=DEV, requires_grad=True)
        F.interpolate(xr, size=(hq, wq), mode="nearest").backward(g.float().permute(0, 3, 1, 2))
        assert rel(dx, xr.grad.permute(0, 2, 3, 1)) < KERNEL_TOL
    a = rnd(1000, 64, seed=3).to(bf16)
    b = rnd(1000, 64, seed=4).to(bf16)
    want = (a.float() + b.float()).to(bf16)
    ops.add_bf16(a, b)
    assert torch.equal(a, want)


def test_mse_grad():
    from audioldm_with_lora_b200 import ops
    nb, hw = 3, 500
    pred = rnd(nb, hw, 8, seed=1)
    noise = rnd(nb, 8, hw, seed=2)
    loss = torch.zeros(1, dtype=torch.float32, device=DEV)
    deps = torch.full((nb, hw, 64), float("nan"), dtype=bf16, device=DEV)
    n = nb * hw * 8
    ops.mse_grad(pred, noise, nb, hw, 64, 1.0 / n, loss, deps)
    diff = pred - noise.permute(0, 2, 1)
    assert abs(loss.item() / n - (diff ** 2).mean().item()) < 1e-5
    assert rel(deps[..., :8], 2 * diff / n) < KERNEL_TOL
    assert (deps[..., 8:] == 0).all()


def test_lora_refresh():
    import numpy as np
    from audioldm_with_lora_b200 import ops
    flat = rnd(1000, seed=1)
    dst = torch.zeros(64, 320, dtype=bf16, device=DEV)
    descs = np.zeros(2, dtype=ops.REFRESH_DTYPE)
    # A [8, 40] at offset 100 -> rows 8..16, cols 0..40 ; B [40, 8] at offset 500 -> transposed into rows 16..24, scaled
    descs[0] = (dst.data_ptr() + 8 * 320 * 2, 100, 320, 8, 40, 0, 1.0, 0)
    descs[1] = (dst.data_ptr() + 16 * 320 * 2, 500, 320, 40, 8, 1, 2.0, 0)
    dd = torch.from_numpy(descs.view(np.uint8)).to(DEV)
    ops.lora_refresh(dd, 2, flat)
    assert torch.equal(dst[8:16, :40], flat[100:420].view(8, 40).to(bf16))
    assert torch.equal(dst[16:24, :40], (2.0 * flat[500:820].view(40, 8).T).to(bf16))
    assert (dst[:8] == 0).all() and (dst[24:] == 0).all() and (dst[8:24, 40:] == 0).all()


@pytest.mark.parametrize("b,s,heads,d", [(2, 256, 8, 32), (2, 1000, 8, 32), (3, 252, 8, 48), (4, 64, 8, 80),
                                          (1, 1024, 8, 32), (2, 188, 8, 96), (1, 300, 4, 64)])
def test_attention_fwd_lse_and_bwd(b, s, heads, d):
    from audioldm_with_lora_b200 import ops
    C = heads * d
    qkv = rnd(b, s, 3 * C, seed=1).to(bf16)
    dout = rnd(b, s, C, seed=2).to(bf16)
    out = torch.full((b, s, C), float("nan"), dtype=bf16, device=DEV)
    lse = torch.empty(b, heads, s, dtype=torch.float32, device=DEV)
    ops.attention_lse(qkv, out, lse, b, s, heads, d)
    qr = qkv.float().requires_grad_(True)
    q, k, v = [t.view(b, s, heads, d).transpose(1, 2) for t in qr.chunk(3, -1)]
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s, C)
    assert rel(out, o.detach()) < 1e-2
    # lse: log2-domain log-sum-exp of the scaled scores
    sc = (q.detach() @ k.detach().transpose(-1, -2)) * d ** -0.5
    lse_ref = torch.logsumexp(sc, -1) * 1.4426950408889634
    assert (lse - lse_ref).abs().max().item() < 2e-2
    o.backward(dout.float())
    dqkv = torch.full((b, s, 3 * C), float("nan"), dtype=bf16, device=DEV)
    delta = torch.empty(b, heads, s, dtype=torch.float32, device=DEV)
    ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, b, s, heads, d)
    assert not torch.isnan(dqkv.float()).any()
    g = qr.grad
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        e = rel(dqkv[..., sl], g[..., sl])
        assert e < ATTN_BWD_TOL, f"{name}: rel-L2 {e:.3e}"
