"""ORACLE (test infrastructure, not product code) -- parity unpinned by the reference.

Torch-eager fp32 restatement of the reference's LoRA fine-tuning step body
(/root/reference/script/train/train_audioldm_lora.py:499-565) with the data pipeline replaced by synthetic latents /
CLAP embeddings as BASELINE.json config 4 states:

    noisy = DDIMScheduler.add_noise(latents, noise, t)                     :504
    pred  = unet(noisy, t, None, class_labels=embeds, scale=1.0)           :539-546   (oracle/unet_ref.py, peft LoRA unmerged)
    loss  = F.mse_loss(pred.float(), noise.float(), reduction="mean")      :549
    loss.backward()                                                         :557   (torch autograd; DDP averages grads)
    torch.optim.AdamW(lora params, lr, betas, weight_decay, eps).step()    :394-403, :563
    get_scheduler("polynomial", num_warmup_steps=0)                         :438-443, :564

Only the LoRA matrices require grad (train:373-376, 394); gradient clipping is a no-op in the reference
 and is therefore absent.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .ddim_ref import DDIMRef
from .unet_ref import LoraSet, UNetSpec, unet_forward

Tensor = torch.Tensor


class TrainRef:
    """adapters: {path: (A [r, in], B [out, r], alpha)}; A and B become leaf tensors that require grad."""

    def __init__(self, sd: Dict[str, Tensor], spec: UNetSpec, adapters: Dict[str, Tuple[Tensor, Tensor, float]],
                 lr: float = 1.0e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-5,
                 num_training_steps: Optional[int] = None, lora_scale: float = 1.0):
        self.sd, self.spec = sd, spec
        self.params: Dict[str, Tuple[Tensor, Tensor, float]] = {}
        leaves: List[Tensor] = []
        for p, (A, B, alpha) in adapters.items():
            A = A.detach().clone().float().requires_grad_(True)
            B = B.detach().clone().float().requires_grad_(True)
            self.params[p] = (A, B, alpha)
            leaves += [A, B]
        self.lora = LoraSet(self.params, scale=lora_scale)
        self.opt = torch.optim.AdamW(leaves, lr=lr, betas=betas, weight_decay=weight_decay, eps=eps)
        self.sched = None
        if num_training_steps is not None:
            lr_end, power = 1e-7, 1.0

            def lam(step: int) -> float:            # diffusers get_polynomial_decay_schedule_with_warmup, no warm-up
                if step > num_training_steps:
                    return lr_end / lr
                return ((lr - lr_end) * (1 - step / num_training_steps) ** power + lr_end) / lr

            self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lam)
        self.noise_sched = DDIMRef()

    def loss_and_grads(self, latents: Tensor, noise: Tensor, timesteps: Tensor, prompt_embeds: Tensor) -> Tensor:
        self.opt.zero_grad(set_to_none=True)
        noisy = self.noise_sched.add_noise(latents.float(), noise.float(), timesteps.long())
        pred = unet_forward(self.sd, self.spec, noisy, timesteps, prompt_embeds.float(), lora=self.lora)
        loss = F.mse_loss(pred.float(), noise.float(), reduction="mean")
        loss.backward()
        return loss.detach()

    def grads(self) -> Dict[str, Tuple[Tensor, Tensor]]:
        return {p: (A.grad.clone(), B.grad.clone()) for p, (A, B, _) in self.params.items()}

    def average_grads_with(self, others: List["TrainRef"]) -> None:
        """DDP semantics: every replica ends up with the mean gradient over the data-parallel group."""
        for p, (A, B, _) in self.params.items():
            for t_idx in range(2):
                ts = [(self.params[p][t_idx])] + [o.params[p][t_idx] for o in others]
                mean = torch.stack([t.grad for t in ts]).mean(0)
                for t in ts:
                    t.grad = mean.clone()

    def optimizer_step(self) -> None:
        self.opt.step()
        if self.sched is not None:
            self.sched.step()

    def train_step(self, latents: Tensor, noise: Tensor, timesteps: Tensor, prompt_embeds: Tensor) -> Tensor:
        loss = self.loss_and_grads(latents, noise, timesteps, prompt_embeds)
        self.optimizer_step()
        return loss
