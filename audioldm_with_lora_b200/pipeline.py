"""`AudioLDMPipeline` with the reference's `__call__` signature, on the B200 engine.

Restates diffusers 0.32.2 `AudioLDMPipeline.__call__` semantics as the reference
drives it: /root/reference/app.py:14 (`pipe(prompt, num_inference_steps=200, audio_length_in_s=10.0)`),
/root/reference/script/inference/generate_audio.py:47-52 (50 steps, 10 s, guidance 5.0),
/root/reference/script/train/train_audioldm_lora.py:142,161 (validation).

The denoising loop -- CFG-duplicated UNet forward, guidance combine, scheduler update -- is ONE
CUDA graph per (batch, length) replayed `num_inference_steps` times: the timestep, the update
coefficients and the next UNet input all live on the device (csrc/sampler.cu), so the loop has no
host synchronisation.  The tail (VAE decode, HiFi-GAN vocoder) runs on the same kernels by default (vae.py, vocoder.py);
b200_vae=False / b200_vocoder=False keep the torch-eager reference path BASELINE.json's north_star describes.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .engine import LATENT_C_PAD
from .model import UNet2DConditionModel
from .scheduler import DDIMScheduler

Tensor = torch.Tensor


@dataclass
class AudioPipelineOutput:
    audios: Union[np.ndarray, Tensor]


class _LoopState:
    """Device buffers + captured graph of one denoising step for a fixed (nb_lat, H, W, do_cfg, steps)."""

    def __init__(self, pipe: "AudioLDMPipeline", nb_lat: int, h: int, w: int, do_cfg: bool, hist_slots: int):
        dev = pipe.device
        eng = pipe.unet.engine
        self.nb_lat, self.h, self.w, self.do_cfg = nb_lat, h, w, do_cfg
        self.nb_unet = nb_lat * (2 if do_cfg else 1)
        c = eng.cfg.in_channels
        f32 = dict(dtype=torch.float32, device=dev)
        self.x = torch.zeros(nb_lat, h * w, c, **f32)                       # latent state, NHWC fp32
        self.x_saved = torch.zeros(nb_lat, h * w, c, **f32) if hist_slots else None
        self.hist = torch.zeros(hist_slots, nb_lat, h * w, c, **f32) if hist_slots else None
        self.xin = torch.zeros(self.nb_unet, h * w, LATENT_C_PAD, dtype=torch.bfloat16, device=dev)
        self.eps = torch.zeros(self.nb_unet, h * w, eng.cfg.out_channels, **f32)
        self.labels = torch.zeros(self.nb_unet, eng.cfg.class_in_dim, **f32)
        self.silu_emb = torch.zeros(self.nb_unet, eng.cfg.temb_channels, dtype=torch.bfloat16, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.table = torch.zeros(1024, 8, **f32)
        self.t_steps = torch.zeros(1024, **f32)
        self.guidance = 1.0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_key = None
        # the timestep / class embedding and the batched time_emb_proj GEMM do not depend on the latents: they run on a
        # side stream (a parallel branch of the step graph) next to conv_in and the first GroupNorm
        self.overlap_embed = dev.type == "cuda" and os.environ.get("B200_EMBED_OVERLAP", "1") != "0"
        self.side = torch.cuda.Stream(device=dev) if self.overlap_embed else None
        self.rowvec: Optional[Tensor] = None

    def one_step(self, eng, guidance: float, overrides=None, branches: int = 1) -> None:
        if self.overlap_embed:
            plan = eng._plan(self.nb_unet, self.h, self.w)
            if self.rowvec is None or self.rowvec.shape[1] != plan["temb_total"]:
                self.rowvec = torch.zeros(self.nb_unet, plan["temb_total"], dtype=torch.float32, device=self.xin.device)
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                eng.embed(self.t_steps, self.step, False, self.labels, None, self.silu_emb)
                ops.conv_gemm(plan["W"]["temb"], self.silu_emb, 1, self.nb_unet, 1, self.rowvec, out_ld=plan["temb_total"])
                ready = torch.cuda.Event()
                ready.record(self.side)
            eng.forward_nhwc(self.xin, self.silu_emb, self.nb_unet, self.h, self.w, self.eps, attn_overrides=overrides,
                             mid_branches=branches, rowvec=self.rowvec, rowvec_ready=ready)
        else:
            eng.embed(self.t_steps, self.step, False, self.labels, None, self.silu_emb)
            eng.forward_nhwc(self.xin, self.silu_emb, self.nb_unet, self.h, self.w, self.eps, attn_overrides=overrides,
                             mid_branches=branches)
        ops.sampler_step(self.eps, self.x, self.x_saved, self.hist, self.table, self.step, guidance, self.do_cfg,
                         self.nb_lat, self.h * self.w, eng.cfg.in_channels, LATENT_C_PAD, self.xin)


class AudioLDMPipeline:
    def __init__(self, unet: UNet2DConditionModel, scheduler: Optional[DDIMScheduler] = None, vae=None, vocoder=None,
                 text_encoder=None, tokenizer=None, tail_dtype: torch.dtype = torch.bfloat16, use_cuda_graph: bool = True,
                 branches: Optional[int] = None, b200_vae: Optional[bool] = None, b200_vocoder: Optional[bool] = None,
                 b200_text_encoder: Optional[bool] = None):
        self.unet = unet
        self.scheduler = scheduler or DDIMScheduler()
        self.vae, self.vocoder = vae, vocoder
        self.text_encoder, self.tokenizer = text_encoder, tokenizer
        self.device = unet.b200_device
        self.tail_dtype = tail_dtype
        self.use_cuda_graph = use_cuda_graph
        # concurrent sub-batch chains inside one denoising step (engine.forward_branched).  Measured on B200 at the
        # bench workload (UNet batch 16): 1 chain 5.53 ms/step, 2 chains 6.19, 4 chains 7.6 -- the step is bound by
        # per-tile latency on already-occupied SMs, not by idle SMs, so one chain is the default.
        self.branches = int(os.environ.get("B200_BRANCHES", "1")) if branches is None else int(branches)
        self._loops: Dict[tuple, _LoopState] = {}
        self._tail_graphs: Dict[tuple, Optional[tuple]] = {}
        self.tail_launches = 0
        # VAE decoder: a torch module with diffusers key names is re-hosted on the sm_100a kernels (vae.B200VaeDecoder,
        # SURVEY 8(f) item 1) unless b200_vae=False / B200_VAE=0 keeps it on the torch-eager reference path.
        if b200_vae is None:
            b200_vae = os.environ.get("B200_VAE", "1") != "0"
        if (self.vae is not None and b200_vae and self.device.type == "cuda" and isinstance(self.vae, torch.nn.Module)
                and any(k.startswith("decoder.mid_block") for k in self.vae.state_dict())):
            from .vae import from_torch_decoder
            self.vae = from_torch_decoder(self.vae, self.device)
        if self.vae is not None:
            self.vae = self.vae.to(self.device, tail_dtype).eval()
        # HiFi-GAN vocoder: transformers' SpeechT5HifiGan is re-hosted the same way (vocoder.B200HifiGan)
        # unless b200_vocoder=False / B200_VOCODER=0.
        if b200_vocoder is None:
            b200_vocoder = os.environ.get("B200_VOCODER", "1") != "0"
        if (self.vocoder is not None and b200_vocoder and self.device.type == "cuda" and isinstance(self.vocoder, torch.nn.Module)
                and type(self.vocoder).__name__ == "SpeechT5HifiGan" and not getattr(self.vocoder.config, "normalize_before", False)
                and len(self.vocoder.config.resblock_kernel_sizes) == 3):
            from .vocoder import from_torch_vocoder
            self.vocoder = from_torch_vocoder(self.vocoder, self.device)
        if self.vocoder is not None:
            self.vocoder = self.vocoder.to(self.device, tail_dtype).eval()
        # CLAP text tower: transformers' ClapTextModelWithProjection is re-hosted the same way (clap.B200ClapTextEncoder,
        # SURVEY 8(f) item 4) unless b200_text_encoder=False / B200_TEXT_ENCODER=0.
        if b200_text_encoder is None:
            b200_text_encoder = os.environ.get("B200_TEXT_ENCODER", "1") != "0"
        if (self.text_encoder is not None and b200_text_encoder and self.device.type == "cuda"
                and isinstance(self.text_encoder, torch.nn.Module) and type(self.text_encoder).__name__ == "ClapTextModelWithProjection"):
            from .clap import from_torch_text_encoder
            self.text_encoder = from_torch_text_encoder(self.text_encoder, self.device)
        nblocks = len(self.vae.config.block_out_channels) if self.vae is not None else 3
        self.vae_scale_factor = 2 ** (nblocks - 1)
        self.last_timing: Dict[str, float] = {}

    # ------------------------------------------------------------------ helpers
    @property
    def _vocoder_cfg(self):
        if self.vocoder is not None:
            return self.vocoder.config.upsample_rates, self.vocoder.config.sampling_rate, self.vocoder.config.model_in_dim
        return [5, 4, 2, 2, 2], 16000, 64

    def _encode_prompt(self, prompt, num_waveforms_per_prompt, do_cfg, negative_prompt, prompt_embeds,
                       negative_prompt_embeds) -> Tuple[Tensor, Optional[Tensor]]:
        if prompt_embeds is None:
            if self.text_encoder is None or self.tokenizer is None:
                raise ValueError("no text encoder is loaded (CLAP weights are not available offline): pass "
                                 "`prompt_embeds=` / `negative_prompt_embeds=` (L2-normalised 512-d CLAP embeddings)")
            prompt = [prompt] if isinstance(prompt, str) else list(prompt)
            tok = self.tokenizer(prompt, padding="max_length", max_length=self.tokenizer.model_max_length,
                                 truncation=True, return_tensors="pt")
            with torch.no_grad():
                pe = self.text_encoder(tok.input_ids.to(self.text_encoder.device),
                                       attention_mask=tok.attention_mask.to(self.text_encoder.device)).text_embeds
            prompt_embeds = F.normalize(pe, dim=-1)
        prompt_embeds = prompt_embeds.to(self.device, torch.float32)
        bs, dim = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_waveforms_per_prompt).view(bs * num_waveforms_per_prompt, dim)
        if do_cfg:
            if negative_prompt_embeds is None:
                if self.text_encoder is None:
                    raise ValueError("classifier-free guidance needs `negative_prompt_embeds=` when no text encoder is loaded")
                neg = [""] * bs if negative_prompt is None else ([negative_prompt] if isinstance(negative_prompt, str) else list(negative_prompt))
                tok = self.tokenizer(neg, padding="max_length", max_length=self.tokenizer.model_max_length,
                                     truncation=True, return_tensors="pt")
                with torch.no_grad():
                    ne = self.text_encoder(tok.input_ids.to(self.text_encoder.device),
                                           attention_mask=tok.attention_mask.to(self.text_encoder.device)).text_embeds
                negative_prompt_embeds = F.normalize(ne, dim=-1)
            negative_prompt_embeds = negative_prompt_embeds.to(self.device, torch.float32)
            if negative_prompt_embeds.shape[0] != bs:
                raise ValueError(f"`negative_prompt_embeds` has batch size {negative_prompt_embeds.shape[0]}, but "
                                 f"`prompt_embeds` has batch size {bs}")
            negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_waveforms_per_prompt).view(-1, dim)
        return prompt_embeds, negative_prompt_embeds if do_cfg else None

    def check_inputs(self, prompt, audio_length_in_s, vocoder_upsample_factor, callback_steps, negative_prompt,
                     prompt_embeds, negative_prompt_embeds):
        min_len = vocoder_upsample_factor * self.vae_scale_factor
        if audio_length_in_s < min_len:
            raise ValueError(f"`audio_length_in_s` has to be a positive value greater than or equal to {min_len}, but "
                             f"is {audio_length_in_s}.")
        if callback_steps is None or not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type "
                             f"{type(callback_steps)}.")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError("Cannot forward both `prompt` and `prompt_embeds`. Please make sure to only forward one of the two.")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` undefined.")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError("Cannot forward both `negative_prompt` and `negative_prompt_embeds`.")
        if prompt_embeds is not None and negative_prompt_embeds is not None and prompt_embeds.shape != negative_prompt_embeds.shape:
            raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed directly, "
                             f"but got: `prompt_embeds` {prompt_embeds.shape} != `negative_prompt_embeds` "
                             f"{negative_prompt_embeds.shape}.")

    def prepare_latents(self, batch, channels, height, generator, latents) -> Tensor:
        _, _, model_in_dim = self._vocoder_cfg
        shape = (batch, channels, height // self.vae_scale_factor, model_in_dim // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                             f"effective batch size of {batch}.")
        if latents is None:
            if isinstance(generator, list):
                latents = torch.cat([torch.randn((1,) + shape[1:], generator=g, device=g.device) for g in generator])
            else:
                gdev = generator.device if generator is not None else "cpu"
                latents = torch.randn(shape, generator=generator, device=gdev)
        elif tuple(latents.shape) != shape:
            raise ValueError(f"Unexpected latents shape, got {tuple(latents.shape)}, expected {shape}")
        return latents.to(self.device, torch.float32) * self.scheduler.init_noise_sigma

    # ------------------------------------------------------------------ the denoising loop
    def denoise(self, latents: Tensor, prompt_embeds: Tensor, negative_prompt_embeds: Optional[Tensor],
                num_inference_steps: int, guidance_scale: float, eta: float = 0.0,
                callback: Optional[Callable] = None, callback_steps: int = 1,
                trace: Optional[List[Tensor]] = None, step_range: Optional[Tuple[int, int]] = None) -> Tensor:
        """latents NCHW fp32 [B,8,H,W] (already scaled by init_noise_sigma) -> denoised NCHW fp32.
        step_range=(first, last): run only steps first .. last-1 of the `num_inference_steps` schedule, `latents` being
        the state before step `first` (teacher-forced per-step parity checks; DDIM only -- PLMS carries history)."""
        eng = self.unet.engine
        sched = self.scheduler
        self.unet.sync_adapters()            # validation inside a training run (train_audioldm_lora.py:599) sees the trained LoRA
        do_cfg = negative_prompt_embeds is not None
        nb, c, h, w = latents.shape
        sched.set_timesteps(num_inference_steps)
        table = sched.step_table(eta)
        t_list = sched.unet_timesteps()
        nsteps = len(t_list)
        key = (nb, h, w, do_cfg, sched.hist_slots)
        st = self._loops.get(key)
        if st is None:
            st = self._loops[key] = _LoopState(self, nb, h, w, do_cfg, sched.hist_slots)
        if nsteps > st.table.shape[0]:
            raise ValueError(f"at most {st.table.shape[0]} sampler steps are supported")
        first, last = (0, nsteps) if step_range is None else (int(step_range[0]), int(step_range[1]))
        if not 0 <= first <= last <= nsteps:
            raise ValueError(f"step_range {step_range} outside the {nsteps}-step schedule")
        if first and sched.hist_slots:
            raise ValueError("step_range needs a history-free sampler (DDIM)")
        # ---- stage inputs (host -> device copies happen here, inside the caller's timed region)
        st.table[:nsteps].copy_(table.to(self.device), non_blocking=True)
        st.t_steps[:nsteps].copy_(torch.tensor(t_list, dtype=torch.float32).to(self.device), non_blocking=True)
        st.step.fill_(first)
        st.x.copy_(latents.permute(0, 2, 3, 1).reshape(nb, h * w, c))
        xb = st.x.to(torch.bfloat16)
        st.xin[:nb, :, :c] = xb
        if do_cfg:
            st.xin[nb:, :, :c] = xb
            st.labels.copy_(torch.cat([negative_prompt_embeds, prompt_embeds]))
        else:
            st.labels.copy_(prompt_embeds)
        guidance = float(guidance_scale)
        overrides = self.unet.custom_attn_processors()
        branches = 1 if overrides else eng.effective_branches(st.nb_unet, self.branches)
        eng._plan(st.nb_unet, h, w)          # make sure weights are packed before the graph key is taken
        if branches > 1:
            from .ops import NUM_SMS
            eng._plan(st.nb_unet // branches, h, w, max(1, NUM_SMS // branches))
        graph_key = (guidance, eng.weights_version, eng.arena_token(st.nb_unet, h, w, branches), branches,
                     tuple(sorted(overrides or ())))
        if self.use_cuda_graph and (st.graph is None or st.graph_key != graph_key):
            # warm-up on a side stream (packs weights, sets func attributes), then capture one step
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                saved = (st.x.clone(), st.xin.clone())
                st.one_step(eng, guidance, overrides, branches)
                st.step.fill_(first); st.x.copy_(saved[0]); st.xin.copy_(saved[1])
                if st.hist is not None:
                    st.hist.zero_()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st.one_step(eng, guidance, overrides, branches)
            st.graph, st.graph_key = g, graph_key
            st.launches_per_step = None
        need_host = callback is not None or trace is not None
        for i in range(first, last):
            if self.use_cuda_graph:
                st.graph.replay()
            else:
                st.one_step(eng, guidance, overrides, branches)
            if need_host:
                cur = st.x.view(nb, h, w, c).permute(0, 3, 1, 2).contiguous()
                if trace is not None:
                    trace.append(cur.clone())
                if callback is not None and i % callback_steps == 0:
                    callback(i, t_list[i], cur)
        return st.x.view(nb, h, w, c).permute(0, 3, 1, 2).contiguous()

    # ------------------------------------------------------------------ tail
    def decode_latents(self, latents: Tensor) -> Tensor:
        z = (latents / self.vae.config.scaling_factor).to(self.tail_dtype)
        out = self.vae.decode(z)
        return out.sample if hasattr(out, "sample") else out       # diffusers' AutoencoderKL returns DecoderOutput

    def mel_spectrogram_to_waveform(self, mel: Tensor) -> Tensor:
        if mel.dim() == 4:
            mel = mel.squeeze(1)
        with torch.no_grad():
            wave = self.vocoder(mel.to(self.tail_dtype))
        return wave.cpu().float()

    def latents_to_waveform(self, latents: Tensor) -> Tensor:
        """VAE decode + vocoder on the device (the reference path's torch kernels, unchanged), replayed from a CUDA
        graph per latent shape: the tail is ~1800 small eager launches, i.e. launch-bound.  Falls back to eager
        launches if a foreign VAE / vocoder cannot be captured.  Returns the device waveform [B, samples]."""
        key = (tuple(latents.shape), self.tail_dtype)
        ent = self._tail_graphs.get(key) if self.use_cuda_graph and latents.is_cuda else None
        if ent is None and self.use_cuda_graph and latents.is_cuda and key not in self._tail_graphs:
            try:
                z = torch.zeros_like(latents, dtype=torch.float32)
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s), torch.no_grad():
                    for _ in range(2):                  # warm-up: cuDNN algorithm selection happens outside capture
                        self.vocoder(self.decode_latents(z).squeeze(1).to(self.tail_dtype))
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                from . import _lib
                n0 = _lib.launch_count
                with torch.cuda.graph(g), torch.no_grad():
                    out = self.vocoder(self.decode_latents(z).squeeze(1).to(self.tail_dtype))
                self.tail_launches = _lib.launch_count - n0        # sm_100a kernel launches captured into the tail graph
                ent = self._tail_graphs[key] = (g, z, out)
            except Exception:                           # noqa: BLE001 -- the tail is not the B200 path: eager is equivalent
                torch.cuda.synchronize()
                ent = self._tail_graphs[key] = None
        if ent is None:
            with torch.no_grad():
                mel = self.decode_latents(latents)
                return self.vocoder(mel.squeeze(1).to(self.tail_dtype))
        g, z, out = ent
        z.copy_(latents)
        g.replay()
        return out

    # ------------------------------------------------------------------ __call__
    @torch.no_grad()
    def __call__(self, prompt=None, audio_length_in_s: Optional[float] = None, num_inference_steps: int = 10,
                 guidance_scale: float = 2.5, negative_prompt=None, num_waveforms_per_prompt: Optional[int] = 1,
                 eta: float = 0.0, generator=None, latents: Optional[Tensor] = None,
                 prompt_embeds: Optional[Tensor] = None, negative_prompt_embeds: Optional[Tensor] = None,
                 return_dict: bool = True, callback: Optional[Callable] = None, callback_steps: Optional[int] = 1,
                 cross_attention_kwargs: Optional[dict] = None, output_type: Optional[str] = "np"):
        rates, sr, model_in_dim = self._vocoder_cfg
        vocoder_upsample_factor = float(np.prod(rates)) / sr
        if audio_length_in_s is None:
            audio_length_in_s = self.unet.config.sample_size * self.vae_scale_factor * vocoder_upsample_factor
        height = int(audio_length_in_s / vocoder_upsample_factor)
        original_waveform_length = int(audio_length_in_s * sr)
        if height % self.vae_scale_factor != 0:
            height = int(np.ceil(height / self.vae_scale_factor)) * self.vae_scale_factor
        self.check_inputs(prompt, audio_length_in_s, vocoder_upsample_factor, callback_steps, negative_prompt,
                          prompt_embeds, negative_prompt_embeds)
        do_cfg = guidance_scale > 1.0
        pe, ne = self._encode_prompt(prompt, num_waveforms_per_prompt, do_cfg, negative_prompt, prompt_embeds,
                                     negative_prompt_embeds)
        batch = pe.shape[0]
        lat = self.prepare_latents(batch, self.unet.config.in_channels, height, generator, latents)
        self.unet.sync_adapters()
        self.unet.engine.set_lora_scale(float((cross_attention_kwargs or {}).get("scale", 1.0)))
        lat = self.denoise(lat, pe, ne, num_inference_steps, guidance_scale, eta, callback, callback_steps)
        if output_type == "latent":
            return AudioPipelineOutput(audios=lat) if return_dict else (lat,)
        if self.vae is None or self.vocoder is None:
            raise ValueError("pipeline was built without vae/vocoder: use output_type='latent'")
        audio = self.latents_to_waveform(lat).cpu().float()[:, :original_waveform_length]
        if output_type == "np":
            audio = audio.numpy()
        elif output_type != "pt":
            raise ValueError(f"unknown output_type {output_type!r}")
        return AudioPipelineOutput(audios=audio) if return_dict else (audio,)
