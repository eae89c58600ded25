"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CLIP text encoder the reference calls once per prompt batch before its
loops (`model.text_encoder(input_ids)[0]`: models/p2p/inversion.py:42,50,296,304, models/p2p/p2p_guidance_forward.py:43,49,
86,92; `clip(ids).last_hidden_state`: models/edict/edict_functions.py:818-838).

The algorithm lives in a third-party dependency that is absent from /root/reference: `transformers` (pinned 4.19.2 in
environment/edict_requirements.txt, unpinned in p2p_requirements.txt / masactrl_requirements.txt), class CLIPTextModel with
the configuration of the SD-1.x `text_encoder/` (openai/clip-vit-large-patch14 text tower).  Published algorithm
(modeling_clip.py: CLIPTextEmbeddings, CLIPAttention, CLIPMLP, CLIPEncoderLayer, CLIPTextTransformer):

    x = token_embedding[ids] + position_embedding[0..76]
    12 x:  x = x + out_proj(softmax((q_proj(h) / 8) k_proj(h)^T + causal_mask) v_proj(h)),  h = LayerNorm1(x)   (12 heads of 64)
           x = x + fc2(quick_gelu(fc1(LayerNorm2(x)))),  quick_gelu(u) = u * sigmoid(1.702 u)
    last_hidden_state = final_layer_norm(x)                       (LayerNorm eps 1e-5; no padding mask is applied)

PARITY PINNED: tests/golden/clip_text.npz holds the output of the installed transformers CLIPTextModel itself
(oracle/make_golden.py clip) on the synthetic weights of pnpinversion_b200/synth.py; tests/test_oracle_cpu.py checks this
file against it and, where transformers is importable, against a live CLIPTextModel.  Only tests/ may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

HEADS = 12


class ClipTextRef:
    def __init__(self, state_dict, dtype=torch.float64):
        self.p = {k: v.to(dtype) for k, v in state_dict.items() if v.is_floating_point()}
        self.dtype = dtype
        self.layers = 0
        while f"text_model.encoder.layers.{self.layers}.layer_norm1.weight" in self.p:
            self.layers += 1

    def _lin(self, x, name):
        return F.linear(x, self.p[name + ".weight"], self.p[name + ".bias"])

    def _ln(self, x, name):
        return F.layer_norm(x, (x.shape[-1],), self.p[name + ".weight"], self.p[name + ".bias"], eps=1e-5)

    def _attn(self, h, p):
        b, n, c = h.shape
        d = c // HEADS

        def heads(t):
            return t.reshape(b, n, HEADS, d).transpose(1, 2)

        q = heads(self._lin(h, p + ".q_proj") * d ** -0.5)
        k = heads(self._lin(h, p + ".k_proj"))
        v = heads(self._lin(h, p + ".v_proj"))
        s = q @ k.transpose(-1, -2)
        mask = torch.full((n, n), float("-inf"), dtype=s.dtype).triu(1)  # token i sees tokens 0..i
        o = torch.softmax(s + mask, dim=-1) @ v
        return self._lin(o.transpose(1, 2).reshape(b, n, c), p + ".out_proj")

    def __call__(self, input_ids):
        """input_ids (B,77) int64 -> last_hidden_state (B,77,768)"""
        t = "text_model."
        ids = torch.as_tensor(input_ids).long()
        x = self.p[t + "embeddings.token_embedding.weight"][ids] + self.p[t + "embeddings.position_embedding.weight"][: ids.shape[1]]
        for i in range(self.layers):
            p = f"{t}encoder.layers.{i}"
            x = x + self._attn(self._ln(x, p + ".layer_norm1"), p + ".self_attn")
            u = self._lin(self._ln(x, p + ".layer_norm2"), p + ".mlp.fc1")
            x = x + self._lin(u * torch.sigmoid(1.702 * u), p + ".mlp.fc2")
        return self._ln(x, t + "final_layer_norm")
