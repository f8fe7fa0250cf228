#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py -q -m gpu > gpurun_out/r02_tests7.log 2>&1; tail -30 gpurun_out/r02_tests7.log
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from audioldm_with_lora_b200 import tail
from audioldm_with_lora_b200.vae import from_torch_decoder
vae = tail.random_vae_decoder(7)
dec = from_torch_decoder(vae, "cuda")
tv = vae.to("cuda", torch.bfloat16).eval()
z = torch.randn(8, 8, 250, 16, device="cuda")
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print("b200 vae decode, 8 clips: %.2f ms (eager launches)" % t(lambda: dec.decode(z)))
with torch.no_grad():
    print("torch bf16 vae decode, 8 clips: %.2f ms" % t(lambda: tv.decode(z.to(torch.bfloat16))))
voc = tail.build_vocoder(0).to("cuda", torch.bfloat16)
mel = torch.randn(8, 1000, 64, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    print("torch bf16 vocoder, 8 clips: %.2f ms" % t(lambda: voc(mel)))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    dec.decode(z)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = dec.decode(z)
print("b200 vae decode, graph replay: %.2f ms" % t(lambda: g.replay()))
PY
