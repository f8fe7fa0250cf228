#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py -q -m gpu -x > gpurun_out/r02_tests33.log 2>&1; tail -12 gpurun_out/r02_tests33.log
timeout 300 python - <<'PY' > gpurun_out/r02_enc_time.log 2>&1
import torch, sys
sys.path.insert(0, '.')
from audioldm_with_lora_b200 import synthetic
from audioldm_with_lora_b200.vae import from_torch_encoder
from oracle import vae_ref
sd = synthetic.random_state_dict_from_shapes(vae_ref.vae_encoder_param_shapes(), seed=11, std=0.05)
enc = from_torch_encoder(sd, "cuda")
mel = torch.randn(8, 1, 1024, 64, device="cuda")
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print("b200 vae encode, 8 clips of 10.24 s: %.2f ms" % t(lambda: enc.encode(mel)))
dsd = {k: v.to("cuda", torch.bfloat16) for k, v in sd.items()}
with torch.no_grad():
    print("torch bf16 (oracle functions) vae encode, 8 clips: %.2f ms" % t(lambda: vae_ref.vae_encode(dsd, mel.to(torch.bfloat16))))
PY
cat gpurun_out/r02_enc_time.log
