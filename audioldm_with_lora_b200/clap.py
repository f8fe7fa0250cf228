"""CLAP text tower (transformers `ClapTextModelWithProjection`) on the sm_100a kernels.

The reference loads it at /root/reference/script/train/train_audioldm_lora.py:368-369 and calls it once per training batch
(`:513-524`: `F.normalize(text_encoder(ids, mask).text_embeds)`), and `AudioLDMPipeline._encode_prompt` calls it once per
pipeline call (/root/reference/app.py:14, generate_audio.py:47-52) -- its 512-d output is the `class_labels` of every
UNet call.  Architecture (cvssp/audioldm-s-full-v2 text_encoder/config.json = RoBERTa-base + projection): token / position /
type embeddings + LayerNorm; 12 post-LN layers {q / k / v (768 -> 768, bias), 12-head attention (head_dim 64), output dense
+ residual + LayerNorm, 768 -> 3072 GELU, 3072 -> 768 + residual + LayerNorm}; pooler = tanh(dense(h[:, 0])); projection
linear1 (768 -> 512), ReLU, linear2 (512 -> 512).

How it maps on the kernels:
  * the padded [B, L] batch is PACKED: only the tokens the attention mask keeps are embedded, as one [sum(len_i), 768] token
    matrix (every layer but the attention is row-wise).  RoBERTa pads on the right, a padded key is masked for every query,
    and only row 0 of each sample reaches the pooler, so dropping the padded rows is exact -- and 512-token padding
    (`datasets.py:128-134`) costs nothing;
  * embeddings + LayerNorm: `ops.embed_layernorm` (position ids = RoBERTa's `padding_idx + cumsum(mask)`);
  * every nn.Linear: `ops.conv1d` with one tap (the tcgen05 implicit-GEMM kernel; bias, residual add, exact-erf GELU / tanh /
    ReLU in its epilogue); q / k / v are one GEMM writing the fused [tokens, 3 x 768] layout the attention kernel reads;
  * attention: `ops.attention` per sample over its own length (head_dim 64, no mask needed once the padding is gone);
  * LayerNorm: `ops.layernorm` (eps 1e-12).

`B200ClapTextEncoder` has the call surface the pipeline and the training loop use: `enc(input_ids, attention_mask=mask)
.text_embeds`, `.device`, `.config`.  The torch module stays the parity partner: it IS the reference's own encoder code.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import ops, packing

Tensor = torch.Tensor


class B200ClapTextEncoder:
    def __init__(self, text_encoder, device="cuda"):
        cfg = text_encoder.config
        if getattr(cfg, "hidden_act", "gelu") != "gelu" or getattr(cfg, "projection_hidden_act", "relu") != "relu":
            raise NotImplementedError("CLAP text tower with hidden_act != gelu or projection_hidden_act != relu")
        if getattr(cfg, "position_embedding_type", "absolute") != "absolute":
            raise NotImplementedError("relative position embeddings")
        self.config = cfg
        self.device = torch.device(device)
        self.c = int(cfg.hidden_size)
        self.heads = int(cfg.num_attention_heads)
        self.d = self.c // self.heads
        self.ff = int(cfg.intermediate_size)
        self.layers = int(cfg.num_hidden_layers)
        self.proj = int(cfg.projection_dim)
        self.eps = float(cfg.layer_norm_eps)
        self.pad = int(cfg.pad_token_id)
        if self.c % 64 or self.ff % 64 or self.proj % 64 or self.d not in (32, 48, 64, 80, 96, 160):
            raise NotImplementedError(f"hidden {self.c} / heads {self.heads}: the kernels need 64-channel blocks and a "
                                      "head_dim of 32 / 48 / 64 / 80 / 96 / 160")
        sd = {k: v.detach().float().cpu() for k, v in text_encoder.state_dict().items()}
        dev = self.device
        f32 = lambda t: t.to(dev, torch.float32).contiguous()      # noqa: E731
        e = "text_model.embeddings."
        self.word, self.pos = f32(sd[e + "word_embeddings.weight"]), f32(sd[e + "position_embeddings.weight"])
        self.type0 = f32(sd[e + "token_type_embeddings.weight"][0])
        self.emb_g, self.emb_b = f32(sd[e + "LayerNorm.weight"]), f32(sd[e + "LayerNorm.bias"])
        self.sd = sd
        self.ln: List[Dict[str, Tensor]] = []
        for i in range(self.layers):
            p = f"text_model.encoder.layer.{i}."
            self.ln.append({"g1": f32(sd[p + "attention.output.LayerNorm.weight"]), "b1": f32(sd[p + "attention.output.LayerNorm.bias"]),
                            "g2": f32(sd[p + "output.LayerNorm.weight"]), "b2": f32(sd[p + "output.LayerNorm.bias"])})
        self._plans: Dict[int, dict] = {}

    # torch-module surface
    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ weights (packed per m-tile count: the tiling depends on it)
    def _pw(self, w: Tensor, b: Tensor, m_tiles: int):
        n, k = w.shape
        bn = ops.choose_tiling(n, m_tiles, k // 64, allow_split=False)[0]
        return packing.pack([w], b, bn, 1, k, device=self.device)

    def _plan(self, m_tiles: int) -> dict:
        if m_tiles in self._plans:
            return self._plans[m_tiles]
        sd, W = self.sd, {}
        for i in range(self.layers):
            p = f"text_model.encoder.layer.{i}."
            a = p + "attention.self."
            W[f"{i}.qkv"] = self._pw(torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]]),
                                     torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]]), m_tiles)
            W[f"{i}.out"] = self._pw(sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"], m_tiles)
            W[f"{i}.ff1"] = self._pw(sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"], m_tiles)
            W[f"{i}.ff2"] = self._pw(sd[p + "output.dense.weight"], sd[p + "output.dense.bias"], m_tiles)
        W["pooler"] = self._pw(sd["text_model.pooler.dense.weight"], sd["text_model.pooler.dense.bias"], 1)
        W["proj1"] = self._pw(sd["text_projection.linear1.weight"], sd["text_projection.linear1.bias"], 1)
        W["proj2"] = self._pw(sd["text_projection.linear2.weight"], sd["text_projection.linear2.bias"], 1)
        self._plans[m_tiles] = W
        return W

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, input_ids: Tensor, attention_mask: Optional[Tensor] = None, **_):
        ids = torch.as_tensor(input_ids).cpu().long()
        if ids.dim() != 2:
            raise ValueError("input_ids must be [batch, tokens]")
        nb, L = ids.shape
        mask = torch.ones_like(ids) if attention_mask is None else torch.as_tensor(attention_mask).cpu().long()
        if mask.shape != ids.shape:
            raise ValueError("attention_mask must have the shape of input_ids")
        lens = mask.sum(1)
        if not bool((mask == (torch.arange(L)[None] < lens[:, None]).long()).all()) or int(lens.min()) < 1:
            raise NotImplementedError("attention masks must be right-padded with at least one kept token per prompt "
                                      "(what RobertaTokenizerFast produces)")
        # RoBERTa position ids over the ORIGINAL ids (create_position_ids_from_input_ids: pad tokens do not count)
        notpad = (ids != self.pad).long()
        pos = torch.cumsum(notpad, 1) * notpad + self.pad
        keep = mask.bool()
        dev, c = self.device, self.c
        ids_p = ids[keep].to(dev, torch.int32)
        pos_p = pos[keep].to(dev, torch.int32)
        m = int(ids_p.numel())
        starts = [0] + torch.cumsum(lens, 0).tolist()
        W = self._plan(max(1, math.ceil(m / 128)))
        bf = dict(dtype=torch.bfloat16, device=dev)
        x = torch.empty(m, c, **bf)
        ops.embed_layernorm(ids_p, pos_p, self.word, self.pos, self.type0, self.emb_g, self.emb_b, self.eps, x)
        qkv, att, y = torch.empty(m, 3 * c, **bf), torch.empty(m, c, **bf), torch.empty(m, c, **bf)
        hid = torch.empty(m, self.ff, **bf)
        lin = dict(dh0=0, dh_step=1)
        # samples of equal length share one attention launch when they are neighbours (a padded batch of equal prompts)
        groups, i = [], 0
        while i < nb:
            j = i
            while j + 1 < nb and int(lens[j + 1]) == int(lens[i]):
                j += 1
            groups.append((starts[i], j - i + 1, int(lens[i])))
            i = j + 1
        for li in range(self.layers):
            ops.conv1d(W[f"{li}.qkv"], x, 1, m, qkv, **lin)
            for row0, cnt, ln in groups:
                ops.attention(qkv[row0: row0 + cnt * ln], att[row0: row0 + cnt * ln], cnt, ln, self.heads, self.d)
            ops.conv1d(W[f"{li}.out"], att, 1, m, y, residual=x, **lin)
            ops.layernorm(y, m, c, self.ln[li]["g1"], self.ln[li]["b1"], self.eps, x)
            ops.conv1d(W[f"{li}.ff1"], x, 1, m, hid, act_tanh=2, **lin)
            ops.conv1d(W[f"{li}.ff2"], hid, 1, m, y, residual=x, **lin)
            ops.layernorm(y, m, c, self.ln[li]["g2"], self.ln[li]["b2"], self.eps, x)
        first = x[torch.tensor(starts[:-1], device=dev)].contiguous()                       # [nb, c]: the <s> token of each prompt
        pooled = torch.empty(nb, c, **bf)
        ops.conv1d(W["pooler"], first, 1, nb, pooled, act_tanh=True, **lin)
        h1 = torch.empty(nb, self.proj, **bf)
        ops.conv1d(W["proj1"], pooled, 1, nb, h1, act_slope=0.0, **lin)
        out = torch.empty(nb, self.proj, dtype=torch.float32, device=dev)
        ops.conv1d(W["proj2"], h1, 1, nb, out, **lin)
        return SimpleNamespace(text_embeds=out, last_hidden_state=None, pooler_output=pooled.float())


def from_torch_text_encoder(text_encoder, device="cuda") -> B200ClapTextEncoder:
    """transformers `ClapTextModelWithProjection` -> the sm_100a implementation (same weights)."""
    return B200ClapTextEncoder(text_encoder, device)


def build_text_encoder(layers: int = 12, seed: int = 0, scale: float = 3.0):
    """Random-init transformers `ClapTextModelWithProjection` with the cvssp/audioldm-s-full-v2 text_encoder config (no
    checkpoint offline): RoBERTa-base geometry, `layers` encoder layers.  Weights are scaled up from transformers' 0.02-std
    init so that attention and the LayerNorms see O(1) activations (a parity test on near-zero scores would be vacuous)."""
    from transformers import ClapTextConfig, ClapTextModelWithProjection
    cfg = ClapTextConfig(num_hidden_layers=layers, projection_dim=512)
    torch.manual_seed(seed)
    enc = ClapTextModelWithProjection(cfg).eval()
    with torch.no_grad():
        for name, p in enc.named_parameters():
            if p.dim() >= 2:
                p.mul_(scale)
            elif name.endswith("bias"):
                p.normal_(0.0, 0.05)
    return enc
