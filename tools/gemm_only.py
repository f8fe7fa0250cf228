#!/usr/bin/env python
"""One conv / linear layer shape of the step in a loop (timing A/B of tilings, and the target of ncu --set full captures).
   python tools/gemm_only.py M N K [taps=1] [geglu=0] [bn=auto] [pair=auto] [reps=50]
M = rows (taps=9: nb=16 images of M/16 pixels, width from the UNet level), N = stored output columns (geglu: N gated outputs =
2N GEMM columns), K = taps * channels."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops, packing  # noqa: E402

a = sys.argv[1:]
M, N, K = int(a[0]), int(a[1]), int(a[2])
taps = int(a[3]) if len(a) > 3 else 1
geglu = bool(int(a[4])) if len(a) > 4 else False
bn = int(a[5]) if len(a) > 5 and a[5] != "auto" else None
pair = None if len(a) <= 6 or a[6] == "auto" else bool(int(a[6]))
reps = int(a[7]) if len(a) > 7 else 50
extra = a[8] if len(a) > 8 else ""            # any of: r (residual), v (per-image row vector = the temb add)
c = K // taps
if taps == 9:
    nb = 16
    hw = M // nb
    w = {4000: 16, 1000: 8, 252: 4, 64: 2}[hw]
    h = hw // w
else:
    nb, h, w = 1, M, 1
n_gemm = 2 * N if geglu else N
wt = torch.randn(n_gemm, K) * K ** -0.5
mt = ops.num_m_tiles(nb, h, w)
if bn is None:
    bn = ops.choose_tiling(n_gemm, mt, K // 64, geglu=geglu, allow_split=False)[0]
pw = packing.pack([wt], torch.zeros(n_gemm), bn, taps, c, geglu=geglu, device="cuda")
x = torch.randn(nb * h * w, c, device="cuda").to(torch.bfloat16)
out = torch.empty(nb * h * w, N, dtype=torch.bfloat16, device="cuda")
res = torch.randn(nb * h * w, N, device="cuda").to(torch.bfloat16) if "r" in extra else None
rv = torch.randn(nb, N, device="cuda") if "v" in extra else None
kw = dict(cta_pair=pair, residual=res, rowvec=rv, rowvec_ld=N if rv is not None else 0)
for _ in range(5):
    ops.conv_gemm(pw, x, nb, h, w, out, **kw)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(reps):
        ops.conv_gemm(pw, x, nb, h, w, out, **kw)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"M={M} N={N} K={K} taps={taps} geglu={int(geglu)} bn={bn} pair={pair} extra={extra!r}: {us:.2f} us  {2.0 * M * n_gemm * K / us / 1e6:.0f} TFLOP/s")
