#!/usr/bin/env python
"""Where does the torch tail (VAE decode + HiFi-GAN vocoder, 8 clips of 10 s) spend its time, and do memory formats help?"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import tail  # noqa: E402

dev = "cuda"
vae = tail.random_vae_decoder(7).to(dev, torch.bfloat16).eval()
voc = tail.build_vocoder(0).to(dev, torch.bfloat16).eval()
z = torch.randn(8, 8, 250, 16, device=dev, dtype=torch.bfloat16)


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    mel = vae.decode(z)
    print("vae decode NCHW bf16        : %.1f ms" % t(lambda: vae.decode(z)))
    print("vocoder bf16                : %.1f ms" % t(lambda: voc(mel.squeeze(1))))
    vae_cl = tail.random_vae_decoder(7).to(dev, torch.bfloat16).eval().to(memory_format=torch.channels_last)
    zc = z.contiguous(memory_format=torch.channels_last)
    mel2 = vae_cl.decode(zc)
    print("vae decode channels_last    : %.1f ms   (rel diff %.2e)" % (t(lambda: vae_cl.decode(zc)),
          ((mel2.float() - mel.float()).norm() / mel.float().norm()).item()))
    torch.backends.cudnn.benchmark = True
    print("vae decode NCHW, cudnn.benchmark         : %.1f ms" % t(lambda: vae.decode(z)))
    print("vae decode channels_last, cudnn.benchmark: %.1f ms" % t(lambda: vae_cl.decode(zc)))
    print("vocoder, cudnn.benchmark                 : %.1f ms" % t(lambda: voc(mel.squeeze(1))))
