#!/usr/bin/env python
"""Small invocation of every kernel family (sampling step + training step at a tiny latent) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Shapes are chosen ragged (odd heights, partial tiles) so that out-of-bounds accesses would show."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import audioldm_with_lora_b200 as b2  # noqa: E402
from audioldm_with_lora_b200 import synthetic  # noqa: E402
from audioldm_with_lora_b200.train import LoraTrainer  # noqa: E402

cfg = b2.CONFIGS["S"]
unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device="cuda")
unet.load_state_dict(synthetic.random_lora_state_dict(cfg, 8, fmt="peft"), strict=False)
pipe = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), use_cuda_graph=False)
x = synthetic.initial_latents(2, 25).cuda()
pos, neg = [t.cuda() for t in synthetic.clap_embeddings(2)]
out = pipe.denoise(x, pos, neg, 2, 2.5)
torch.cuda.synchronize()
print("sampling ok", float(out.abs().mean()))
trainer = LoraTrainer(unet, lr=1e-4)
g = torch.Generator().manual_seed(5)
lat, noise = torch.randn(2, 8, 20, 16, generator=g), torch.randn(2, 8, 20, 16, generator=g)
t = torch.randint(0, 1000, (2,), generator=g)
loss = trainer.train_step(lat, noise, t, synthetic.clap_embeddings(2)[0])
torch.cuda.synchronize()
print("training ok", float(loss), float(trainer.flat_g.abs().sum()))
