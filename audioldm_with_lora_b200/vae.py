"""`AutoencoderKL.decode` of the AudioLDM VAE on the sm_100a kernels.

The reference loads the VAE at /root/reference/script/train/train_audioldm_lora.py:370 and runs its decoder inside
`AudioLDMPipeline.__call__` (`decode_latents`, /root/reference/app.py:14): latents [B, 8, H, 16] / scaling_factor ->
log-mel spectrogram [B, 1, 4H, 64].  Architecture (cvssp/audioldm-s-full-v2 vae/config.json):
post_quant_conv 1x1 -> conv_in 3x3 (8 -> 512) -> mid block (ResNet, single-head attention with head_dim 512, ResNet) ->
3 up blocks of 3 ResNets (512, 256, 128 channels; nearest x2 + 3x3 conv between them) -> GroupNorm + SiLU -> conv_out
(128 -> 1).  GroupNorm: 32 groups, eps 1e-6; ResNets without a time embedding.

Same kernels as the UNet (`ops.conv_gemm` implicit GEMM on tcgen05 with K segments for the 1x1 shortcut, `groupnorm_silu`,
`upsample_nearest`), NHWC bf16 activations from one arena, plus:
  * post_quant_conv is FOLDED into conv_in exactly: W'[tap] = W_in[tap] W_pq on the 8 latent channels and
    W_in[tap] b_pq on a ninth input channel that is 1 on every real pixel -- TMA's zero fill outside the image makes the
    folded bias vanish exactly where the reference's zero padding does;
  * the mid-block attention (4000 tokens at a 10 s clip, ONE head of 512 channels) does not fit the fused attention
    kernel's TMEM budget and runs once per clip, so it is three GEMM shapes + a row softmax: Q = n W_q^T + b_q and
    K = n W_k^T (b_k shifts every score of a row equally: softmax-invariant, dropped), S_b = Q_b K_b^T (K_b is the "weight"
    operand as it lies), P_b = softmax(S_b / sqrt(512)) (`ops.softmax_rows`, fp32 scores, bf16 probabilities),
    V_b^T = W_v n_b^T (computed transposed, so P_b V_b is again a plain K-major GEMM), O_b = P_b V_b, and b_v -- every
    row of P sums to one -- folded into to_out's bias.

`B200VaeDecoder` has the `decode(z)` / `.config` surface `AudioLDMPipeline` uses, so it is a drop-in for the torch-eager
`tail.AutoencoderKLDecoder` (which stays the reference-path implementation and the parity partner).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from . import ops, packing
from .engine import Arena

Tensor = torch.Tensor
C_IN_PAD = 64                  # 8 latent channels + the constant-one channel, padded to one 64-channel K block
ONE_CH = 8


class B200VaeDecoder:
    class _Cfg:
        scaling_factor = 0.9227914214134216
        block_out_channels = (128, 256, 512)
        latent_channels = 8

    def __init__(self, state_dict: Dict[str, Tensor], device="cuda", block_out=(128, 256, 512), latent: int = 8,
                 layers: int = 2, groups: int = 32, eps: float = 1e-6):
        self.config = self._Cfg()
        self.config.block_out_channels = tuple(block_out)
        self.device = torch.device(device)
        self.block_out, self.latent, self.layers, self.groups, self.eps = tuple(block_out), latent, layers, groups, eps
        self.sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        self._plans: Dict[Tuple[int, int, int], dict] = {}
        self.arena: Optional[Arena] = None
        self._small = {k: v.to(self.device).contiguous() for k, v in self.sd.items()
                       if ("norm" in k) and v.dim() == 1}
        self._persist: Dict[tuple, Tensor] = {}

    # torch-module surface used by AudioLDMPipeline
    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ weights
    def _pw(self, segs, bias, m_tiles, ntaps, c0, c1=0, block_n=None):
        n = segs[0].shape[0]
        num_kb = (ntaps * c0 + c1) // 64
        bn = block_n or ops.choose_tiling(n, m_tiles, num_kb, allow_split=False)[0]
        return packing.pack(segs, bias, bn, ntaps, c0, c1, device=self.device)

    def _res_names(self):
        top = self.block_out[-1]
        out = [("decoder.mid_block.resnets.0", top, top, 0), ("decoder.mid_block.resnets.1", top, top, 0)]
        prev = top
        for i, c in enumerate(reversed(self.block_out)):
            for j in range(self.layers + 1):
                out.append((f"decoder.up_blocks.{i}.resnets.{j}", prev, c, i))
                prev = c
        return out

    def _plan(self, nb: int, h: int, w: int) -> dict:
        key = (nb, h, w)
        if key in self._plans:
            return self._plans[key]
        sd, W = self.sd, {}
        sizes = [(h << i, w << i) for i in range(len(self.block_out))]

        def tiles(lvl):
            return ops.num_m_tiles(nb, *sizes[lvl])

        # post_quant_conv folded into conv_in (see the module docstring)
        w_in, b_in = sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"]
        w_pq, b_pq = sd["post_quant_conv.weight"][:, :, 0, 0], sd["post_quant_conv.bias"]
        top = w_in.shape[0]
        fold = torch.zeros(top, 3, 3, C_IN_PAD)
        fold[..., : self.latent] = torch.einsum("ockl,ci->okli", w_in, w_pq)
        fold[..., ONE_CH] = torch.einsum("ockl,c->okl", w_in, b_pq)
        W["conv_in"] = self._pw([fold.reshape(top, 9 * C_IN_PAD)], b_in, tiles(0), 9, C_IN_PAD)
        for name, cin, cout, lvl in self._res_names():
            W[name + ".conv1"] = self._pw([packing.conv3x3_to_k(sd[name + ".conv1.weight"])], sd[name + ".conv1.bias"],
                                         tiles(lvl), 9, cin)
            w2 = packing.conv3x3_to_k(sd[name + ".conv2.weight"])
            if cin != cout:
                ws = sd[name + ".conv_shortcut.weight"][:, :, 0, 0]
                W[name + ".conv2"] = self._pw([w2, ws], sd[name + ".conv2.bias"] + sd[name + ".conv_shortcut.bias"],
                                             tiles(lvl), 9, cout, cin)
            else:
                W[name + ".conv2"] = self._pw([w2], sd[name + ".conv2.bias"], tiles(lvl), 9, cout)
        for i in range(len(self.block_out) - 1):
            n = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            W[n] = self._pw([packing.conv3x3_to_k(sd[n + ".weight"])], sd[n + ".bias"], tiles(i + 1), 9, sd[n + ".weight"].shape[1])
        # conv_out: one output channel -> 8 stored columns (column 0 is the mel bin), fp32
        wo = packing.conv3x3_to_k(sd["decoder.conv_out.weight"])
        wo8 = torch.zeros(8, wo.shape[1]); wo8[:1] = wo
        bo8 = torch.zeros(8); bo8[:1] = sd["decoder.conv_out.bias"]
        W["conv_out"] = self._pw([wo8], bo8, tiles(len(self.block_out) - 1), 9, wo.shape[1] // 9, block_n=32)
        plan = {"W": W, "sizes": sizes}
        plan.update(self._pack_mid_attention(W, "decoder.mid_block.attentions.0", top, nb, h * w))
        self._plans[key] = plan
        return plan

    def _pack_mid_attention(self, W: dict, a: str, c: int, nb: int, t: int) -> dict:
        """Weights of the single-head mid-block attention over t tokens per image (see the module docstring)."""
        sd = self.sd
        mt = math.ceil(nb * t / 128)
        W["attn.q"] = self._pw([sd[a + ".to_q.weight"]], sd[a + ".to_q.bias"], mt, 1, c)
        W["attn.k"] = self._pw([sd[a + ".to_k.weight"]], None, mt, 1, c)
        W["attn.wv"] = sd[a + ".to_v.weight"].to(self.device, torch.bfloat16).contiguous()            # the A operand of V^T = W_v n^T
        b_out = sd[a + ".to_out.0.bias"] + sd[a + ".to_out.0.weight"] @ sd[a + ".to_v.bias"]
        W["attn.out"] = self._pw([sd[a + ".to_out.0.weight"]], b_out, mt, 1, c)
        t_pad = (t + 63) // 64 * 64
        return {"t": t, "t_pad": t_pad,
                "bn_s": ops.choose_tiling(t, math.ceil(t / 128), c // 64, allow_split=False)[0],
                "bn_vt": ops.choose_tiling(t, math.ceil(c / 128), c // 64, allow_split=False)[0],
                "bn_o": 256 if c % 256 == 0 else (128 if c % 128 == 0 else 64)}

    def _buf(self, key: tuple, shape, dtype) -> Tensor:
        """Persistent zero-initialised device buffer (padding regions are never written, so they stay zero)."""
        if key not in self._persist:
            self._persist[key] = torch.zeros(shape, dtype=dtype, device=self.device)
        return self._persist[key]

    def _ensure_arena(self, nb: int, h: int, w: int) -> Arena:
        top_px = nb * (h << 2) * (w << 2)
        need = top_px * 256 * 2 * 7 + (128 << 20)
        if self.arena is None or self.arena.buf.numel() < need:
            self.arena = Arena(need, self.device)
        return self.arena

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def decode(self, z: Tensor) -> Tensor:
        """z [B, 8, H, 16] (latents / scaling_factor, any float dtype) -> mel [B, 1, 4H, 64] fp32."""
        nb, cz, h, w = z.shape
        if cz != self.latent:
            raise ValueError(f"expected {self.latent} latent channels, got {cz}")
        if not z.is_cuda:
            raise RuntimeError("B200VaeDecoder needs CUDA tensors (no CPU fallback)")
        plan = self._plan(nb, h, w)
        W, sizes, S = plan["W"], plan["sizes"], self._small
        ar = self._ensure_arena(nb, h, w)
        bf16 = torch.bfloat16

        def M(lvl):
            return nb * sizes[lvl][0] * sizes[lvl][1]

        def gn(x, c, lvl, name, silu, rows_slack=0):
            hh, ww = sizes[lvl]
            y = ar.alloc((M(lvl) + rows_slack, c), bf16)
            ops.groupnorm_silu(x, c, None, 0, nb, hh * ww, S[name + ".weight"], S[name + ".bias"], self.eps, silu, y, self.groups)
            return y

        def conv(name, a0, lvl, *, a1=None, residual=None, out=None):
            pw = W[name]
            hh, ww = sizes[lvl]
            if out is None:
                out = ar.alloc((M(lvl), pw.n_valid), bf16)
            ops.conv_gemm(pw, a0, nb, hh, ww, out, a1=a1, residual=residual)
            return out

        def resnet(name, x, cin, cout, lvl):
            n1 = gn(x, cin, lvl, name + ".norm1", True)
            h1 = conv(name + ".conv1", n1, lvl)
            ar.release(n1)
            n2 = gn(h1, cout, lvl, name + ".norm2", True)
            ar.release(h1)
            out = conv(name + ".conv2", n2, lvl, a1=x) if cin != cout else conv(name + ".conv2", n2, lvl, residual=x)
            ar.release(n2)
            ar.release(x)
            return out

        # ---- input: NCHW -> NHWC, 64 channels (8 latent + the constant-one channel that carries post_quant_conv's bias)
        xin = self._buf(("xin", nb, h, w), (nb, h * w, C_IN_PAD), bf16)
        xin[:, :, ONE_CH] = 1.0
        xin[:, :, : self.latent].copy_(z.permute(0, 2, 3, 1).reshape(nb, h * w, self.latent))
        top = self.block_out[-1]
        x = conv("conv_in", xin, 0)
        x = resnet("decoder.mid_block.resnets.0", x, top, top, 0)
        x = self._mid_attention(plan, ar, x, nb, top, gn)
        x = resnet("decoder.mid_block.resnets.1", x, top, top, 0)
        prev = top
        nlev = len(self.block_out)
        for i, c in enumerate(reversed(self.block_out)):
            for j in range(self.layers + 1):
                x = resnet(f"decoder.up_blocks.{i}.resnets.{j}", x, prev, c, i)
                prev = c
            if i != nlev - 1:
                (hs, ws), (ho, wo) = sizes[i], sizes[i + 1]
                up = ar.alloc((M(i + 1), c), bf16)
                ops.upsample_nearest(x, nb, hs, ws, c, ho, wo, up)
                ar.release(x)
                x = conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", up, i + 1)
                ar.release(up)
        n = gn(x, prev, nlev - 1, "decoder.conv_norm_out", True)
        ar.release(x)
        out8 = self._buf(("out8", nb, h, w), (M(nlev - 1), 8), torch.float32)
        conv("conv_out", n, nlev - 1, out=out8)
        ar.release(n)
        assert not ar.live, f"arena leak: {len(ar.live)} buffers"
        ho, wo = sizes[-1]
        return out8.view(nb, ho, wo, 8)[..., 0].reshape(nb, 1, ho, wo).clone()

    def _mid_attention(self, plan: dict, ar: Arena, x: Tensor, nb: int, c: int, gn, a: str = "decoder.mid_block.attentions.0",
                       lvl: int = 0) -> Tensor:
        W, t, t_pad = plan["W"], plan["t"], plan["t_pad"]
        bf16 = torch.bfloat16
        slack = 256                                            # rows a [n_pad, c] "weight" view may read past the last image
        n = gn(x, c, lvl, a + ".group_norm", False, rows_slack=slack)
        m = nb * t
        q = ar.alloc((m, c), bf16)
        k = ar.alloc((m + slack, c), bf16)
        ops.conv_gemm(W["attn.q"], n[:m], 1, m, 1, q)
        ops.conv_gemm(W["attn.k"], n[:m], 1, m, 1, k[:m])
        s_buf = self._buf(("s", t), (t, t_pad), torch.float32)
        p_buf = self._buf(("p", t), (t, t_pad), bf16)
        vt = self._buf(("vt", t, c), (c, t_pad), bf16)
        o = ar.alloc((m, c), bf16)

        def rows_of(buf, row0, rows, bn):                     # [n_pad, c] view: rows past `rows` only feed discarded columns
            return buf[row0: row0 + (rows + bn - 1) // bn * bn]

        # (ops.gemm_nt, not conv_gemm: both operands were written by earlier kernels of this stream -- conv_gemm's
        #  weight-producer warp does not wait for the previous grid)
        for b in range(nb):
            r0 = b * t
            # S_b = Q_b K_b^T (fp32), the keys of image b as the K-major second operand where they lie
            ops.gemm_nt(q[r0: r0 + t], rows_of(k, r0, t, plan["bn_s"]), t, s_buf, plan["bn_s"], out_ld=t_pad)
            ops.softmax_rows(s_buf, t, t, t_pad, p_buf, scale=c ** -0.5)
            # V_b^T = W_v n_b^T: [c, t] (b_v is folded into to_out's bias)
            ops.gemm_nt(W["attn.wv"], rows_of(n, r0, t, plan["bn_vt"]), t, vt, plan["bn_vt"], out_ld=t_pad)
            # O_b = P_b V_b
            ops.gemm_nt(p_buf, vt, c, o[r0: r0 + t], plan["bn_o"])
        ar.release(q); ar.release(k); ar.release(n)
        out = ar.alloc((m, c), bf16)
        ops.conv_gemm(W["attn.out"], o, 1, m, 1, out, residual=x)
        ar.release(o); ar.release(x)
        return out


def from_torch_decoder(vae, device="cuda") -> B200VaeDecoder:
    """Build the B200 decoder from a torch module with diffusers key names (`tail.AutoencoderKLDecoder`, or diffusers'
    `AutoencoderKL`: only `post_quant_conv.*` and `decoder.*` are read)."""
    sd = {k: v for k, v in vae.state_dict().items() if k.startswith(("post_quant_conv.", "decoder."))}
    cfg = getattr(vae, "config", None)
    block_out = tuple(getattr(cfg, "block_out_channels", (128, 256, 512)))
    dec = B200VaeDecoder(sd, device=device, block_out=block_out)
    if cfg is not None and hasattr(cfg, "scaling_factor"):
        dec.config.scaling_factor = cfg.scaling_factor
    return dec


class DiagonalGaussian:
    """`latent_dist` of `AutoencoderKL.encode` (diffusers DiagonalGaussianDistribution): mean / logvar [B, 8, H, 16] fp32."""

    def __init__(self, mean: Tensor, logvar: Tensor):
        self.mean, self.logvar = mean, logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None) -> Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self) -> Tensor:
        return self.mean


class EncoderOutput:
    def __init__(self, latent_dist: DiagonalGaussian):
        self.latent_dist = latent_dist


class B200VaeEncoder(B200VaeDecoder):
    """`AutoencoderKL.encode` of the AudioLDM VAE on the sm_100a kernels (the training
    loop's `vae.encode(batch["log_mel_spec"]).latent_dist.sample() * scaling_factor`,
    /root/reference/script/train/train_audioldm_lora.py:495-496): log-mel [B, 1, T, 64] -> posterior over latents
    [B, 8, T/4, 16].  conv_in (1 -> 128; the mel bin rides in channel 0 of a 64-channel K block), three down blocks of two
    ResNets (128, 256, 512 channels) with Downsample2D(padding=0) between them (`ops.conv3x3_s2_pad01`: F.pad (0,1,0,1) +
    conv k3 s2 as strided TMA boxes), the mid block (ResNet, head_dim-512 attention as in the decoder, ResNet), GroupNorm +
    SiLU, and conv_out with quant_conv FOLDED in exactly (a 1x1 convolution after a 3x3 one composes: W' = W_q W_out per
    tap, b' = W_q b_out + b_q).  Same kernels / arena as the decoder."""

    def __init__(self, state_dict: Dict[str, Tensor], device="cuda", block_out=(128, 256, 512), latent: int = 8,
                 layers: int = 2, groups: int = 32, eps: float = 1e-6):
        super().__init__(state_dict, device, block_out, latent, layers, groups, eps)

    def _enc_res_names(self):
        out, prev = [], self.block_out[0]
        for i, c in enumerate(self.block_out):
            for j in range(self.layers):
                out.append((f"encoder.down_blocks.{i}.resnets.{j}", prev, c, i))
                prev = c
        top = self.block_out[-1]
        lvl = len(self.block_out) - 1
        out += [("encoder.mid_block.resnets.0", top, top, lvl), ("encoder.mid_block.resnets.1", top, top, lvl)]
        return out

    def _enc_plan(self, nb: int, h: int, w: int) -> dict:
        key = ("enc", nb, h, w)
        if key in self._plans:
            return self._plans[key]
        sd, W = self.sd, {}
        nlev = len(self.block_out)
        sizes = [(h, w)]
        for _ in range(nlev - 1):
            sizes.append(((sizes[-1][0] - 2) // 2 + 1, (sizes[-1][1] - 2) // 2 + 1))

        def tiles(lvl):
            return ops.num_m_tiles(nb, *sizes[lvl])

        w_in = sd["encoder.conv_in.weight"]                                    # [128, 1, 3, 3]
        W["conv_in"] = self._pw([packing.conv3x3_to_k(w_in, C_IN_PAD)], sd["encoder.conv_in.bias"], tiles(0), 9, C_IN_PAD)
        for name, cin, cout, lvl in self._enc_res_names():
            W[name + ".conv1"] = self._pw([packing.conv3x3_to_k(sd[name + ".conv1.weight"])], sd[name + ".conv1.bias"],
                                         tiles(lvl), 9, cin)
            w2 = packing.conv3x3_to_k(sd[name + ".conv2.weight"])
            if cin != cout:
                ws = sd[name + ".conv_shortcut.weight"][:, :, 0, 0]
                W[name + ".conv2"] = self._pw([w2, ws], sd[name + ".conv2.bias"] + sd[name + ".conv_shortcut.bias"],
                                             tiles(lvl), 9, cout, cin)
            else:
                W[name + ".conv2"] = self._pw([w2], sd[name + ".conv2.bias"], tiles(lvl), 9, cout)
        for i in range(nlev - 1):
            n = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            W[n] = self._pw([packing.conv3x3_to_k(sd[n + ".weight"])], sd[n + ".bias"], tiles(i + 1), 9, sd[n + ".weight"].shape[1],
                            split=False)
        # conv_out with quant_conv folded in: 2 * latent fp32 columns (mean | logvar)
        w_out, b_out = sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"]
        w_q, b_q = sd["quant_conv.weight"][:, :, 0, 0], sd["quant_conv.bias"]
        fold = torch.einsum("oc,cikl->oikl", w_q, w_out)
        W["conv_out"] = self._pw([packing.conv3x3_to_k(fold)], w_q @ b_out + b_q, tiles(nlev - 1), 9, w_out.shape[1], block_n=32)
        plan = {"W": W, "sizes": sizes}
        top = self.block_out[-1]
        plan.update(self._pack_mid_attention(W, "encoder.mid_block.attentions.0", top, nb, sizes[-1][0] * sizes[-1][1]))
        self._plans[key] = plan
        return plan

    def _pw(self, segs, bias, m_tiles, ntaps, c0, c1=0, block_n=None, split=True):
        return super()._pw(segs, bias, m_tiles, ntaps, c0, c1, block_n)

    @torch.no_grad()
    def encode(self, mel: Tensor) -> EncoderOutput:
        """mel [B, 1, T, 64] (any float dtype; T and 64 at least 4) -> EncoderOutput(latent_dist) like diffusers."""
        nb, cm, h, w = mel.shape
        if cm != 1:
            raise ValueError(f"expected a 1-channel log-mel spectrogram, got {cm} channels")
        # (no CPU fallback: ops -> _lib.ptr() refuses CPU tensors)
        plan = self._enc_plan(nb, h, w)
        W, sizes, S = plan["W"], plan["sizes"], self._small
        nlev = len(self.block_out)
        need = nb * h * w * self.block_out[0] * 2 * 5 + (128 << 20)
        if self.arena is None or self.arena.buf.numel() < need:
            self.arena = Arena(need, self.device)
        ar = self.arena
        bf16 = torch.bfloat16

        def M(lvl):
            return nb * sizes[lvl][0] * sizes[lvl][1]

        def gn(x, c, lvl, name, silu, rows_slack=0):
            hh, ww = sizes[lvl]
            y = ar.alloc((M(lvl) + rows_slack, c), bf16)
            ops.groupnorm_silu(x, c, None, 0, nb, hh * ww, S[name + ".weight"], S[name + ".bias"], self.eps, silu, y, self.groups)
            return y

        def conv(name, a0, lvl, *, a1=None, residual=None, out=None):
            pw = W[name]
            hh, ww = sizes[lvl]
            if out is None:
                out = ar.alloc((M(lvl), pw.n_valid), bf16)
            ops.conv_gemm(pw, a0, nb, hh, ww, out, a1=a1, residual=residual)
            return out

        def resnet(name, x, cin, cout, lvl):
            n1 = gn(x, cin, lvl, name + ".norm1", True)
            h1 = conv(name + ".conv1", n1, lvl)
            ar.release(n1)
            n2 = gn(h1, cout, lvl, name + ".norm2", True)
            ar.release(h1)
            out = conv(name + ".conv2", n2, lvl, a1=x) if cin != cout else conv(name + ".conv2", n2, lvl, residual=x)
            ar.release(n2)
            ar.release(x)
            return out

        xin = self._buf(("enc_xin", nb, h, w), (nb, h * w, C_IN_PAD), bf16)
        xin[:, :, 0].copy_(mel.reshape(nb, h * w))
        x = conv("conv_in", xin, 0)
        prev = self.block_out[0]
        for i, c in enumerate(self.block_out):
            for j in range(self.layers):
                x = resnet(f"encoder.down_blocks.{i}.resnets.{j}", x, prev, c, i)
                prev = c
            if i != nlev - 1:
                hs, ws = sizes[i]
                y = ar.alloc((M(i + 1), c), bf16)
                ops.conv3x3_s2_pad01(W[f"encoder.down_blocks.{i}.downsamplers.0.conv"], x, nb, hs, ws, y)
                ar.release(x)
                x = y
        top, lvl = self.block_out[-1], nlev - 1
        x = resnet("encoder.mid_block.resnets.0", x, top, top, lvl)
        x = self._mid_attention(plan, ar, x, nb, top, gn, a="encoder.mid_block.attentions.0", lvl=lvl)
        x = resnet("encoder.mid_block.resnets.1", x, top, top, lvl)
        n = gn(x, top, lvl, "encoder.conv_norm_out", True)
        ar.release(x)
        mom = self._buf(("enc_mom", nb, h, w), (M(lvl), 2 * self.latent), torch.float32)
        conv("conv_out", n, lvl, out=mom)
        ar.release(n)
        assert not ar.live, f"arena leak: {len(ar.live)} buffers"
        ho, wo = sizes[-1]
        m4 = mom.view(nb, ho, wo, 2 * self.latent).permute(0, 3, 1, 2)
        return EncoderOutput(DiagonalGaussian(m4[:, : self.latent].contiguous(), m4[:, self.latent:].contiguous()))


def from_torch_encoder(vae, device="cuda") -> B200VaeEncoder:
    """Build the B200 encoder from a module / state dict with diffusers key names (`encoder.*`, `quant_conv.*`)."""
    sd = vae if isinstance(vae, dict) else vae.state_dict()
    sd = {k: v for k, v in sd.items() if k.startswith(("encoder.", "quant_conv."))}
    cfg = getattr(vae, "config", None)
    block_out = tuple(getattr(cfg, "block_out_channels", (128, 256, 512)))
    enc = B200VaeEncoder(sd, device=device, block_out=block_out)
    if cfg is not None and hasattr(cfg, "scaling_factor"):
        enc.config.scaling_factor = cfg.scaling_factor
    return enc
