#!/usr/bin/env python
"""Per-adapter LoRA-gradient error of the B200 training step vs the fp32 oracle (debug aid)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.test_gpu_train import _setup, _batch, _flat, rel  # noqa: E402

rank, nb, h = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
targets = tuple(sys.argv[4].split(",")) if len(sys.argv) > 4 else ("to_q", "to_k", "to_v", "to_out.0")
alpha = float(sys.argv[5]) if len(sys.argv) > 5 else None
unet, trainer, ref = _setup(rank, targets, alpha)
lat, noise, t, emb = _batch(nb, h)
loss_ref = ref.loss_and_grads(lat, noise, t, emb)
noisy = ref.noise_sched.add_noise(lat, noise, t)
trainer.flat_g.zero_()
loss = trainer.forward_backward(noisy.cuda(), t.cuda(), emb.cuda(), noise.cuda())
print("loss", loss.item(), loss_ref.item())
g_ref = _flat(ref, trainer, "grad")
g = trainer.flat_g.cpu()
print("flat rel", rel(g, g_ref), "norm", g_ref.norm().item())
rows = []
for p, s in trainer.slots.items():
    for nm, off in (("A", s.off_a), ("B", s.off_b)):
        sl = slice(off, off + s.r * s.c)
        rows.append((rel(g[sl], g_ref[sl]), g_ref[sl].norm().item(), g[sl].norm().item(), p + "." + nm))
rows.sort(reverse=True)
for r in rows[:24]:
    print(f"{r[0]:.3e}  ref {r[1]:.3e}  got {r[2]:.3e}  {r[3]}")
