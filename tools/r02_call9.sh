#!/bin/bash
# vocoder on the sm_100a kernels: parity tests, then timing against torch bf16 (8 clips of 10 s)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -x > gpurun_out/r02_tests9.log 2>&1; tail -30 gpurun_out/r02_tests9.log
timeout 600 python - <<'PY' > gpurun_out/r02_voc_time.log 2>&1
import torch, sys
sys.path.insert(0, '.')
from audioldm_with_lora_b200 import tail, _lib
from audioldm_with_lora_b200.vocoder import from_torch_vocoder
voc = tail.build_vocoder(0)
mine = from_torch_vocoder(voc, "cuda")
tv = voc.to("cuda", torch.bfloat16).eval()
mel = torch.randn(8, 1000, 64, device="cuda")
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print("b200 vocoder, 8 clips: %.2f ms (eager launches)" % t(lambda: mine(mel)))
with torch.no_grad():
    print("torch bf16 vocoder, 8 clips: %.2f ms" % t(lambda: tv(mel.to(torch.bfloat16))))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    mine(mel)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = mine(mel)
print("b200 vocoder, graph replay: %.2f ms" % t(lambda: g.replay()))
# per-launch profile
_lib.PROFILE = []
mine(mel); torch.cuda.synchronize()
rows = {}
for name, e0, e1, info, args in _lib.PROFILE:
    key = (name, (info or {}).get("m"), (info or {}).get("n"), (info or {}).get("k"), (info or {}).get("bn"))
    r = rows.setdefault(key, [0, 0.0, 0.0]); r[0] += 1; r[1] += e0.elapsed_time(e1) * 1e3; r[2] += (info or {}).get("flops", 0.0)
_lib.PROFILE = None
tot = sum(r[1] for r in rows.values())
print("eager per-launch sum %.1f us" % tot)
for k, r in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %3d x %8.1f us = %9.1f us  %5.1f%%  %6.0f TFLOP/s" % (str(k), r[0], r[1] / r[0], r[1], 100 * r[1] / tot, r[2] / r[1] / 1e6 if r[1] else 0))
PY
cat gpurun_out/r02_voc_time.log
