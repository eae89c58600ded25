"""TEST INFRASTRUCTURE ONLY -- CPU restatement (PyTorch ops on the host, fp64 by default) of the SD-1.x
UNet2DConditionModel forward that the reference drives.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module; it is the *checker*, never the product path (the product is libpnpinv.so and fails loudly without it).

The algorithm lives in the un-vendored dependency `diffusers` (pinned 0.10.0 for P2P, environment/p2p_requirements.txt:1);
the in-tree arithmetic spec followed here is the reference's vendored copy of diffusers 0.3.0,
`/root/reference/models/edict/my_diffusers/` (file:line cited at each function).  Pinning: `tests/test_oracle_cpu.py`
checks this restatement against outputs of the reference's own classes (run in the build container through
`oracle/ref_shim.py`, fixtures committed under `tests/golden/` by `oracle/make_golden.py`).  Parity is therefore
pinned against the reference itself run here; the reference ships no golden vectors of its own.

`attn_hook(probs, is_cross, place_in_unet)` is the reference's controller seam
(`models/p2p/attention_control.py:43-45`): it receives the materialised (B*8, N, K) probabilities, batch-major heads.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

HEADS = 8
BLOCK_OUT = (320, 640, 1280, 1280)


def timestep_embedding(t: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """my_diffusers/models/embeddings.py:21-60 with flip_sin_to_cos=True, downscale_freq_shift=0 (fp64)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float64) / half
    emb = t[:, None].double() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)


class UNetRef:
    def __init__(self, state_dict: Dict[str, torch.Tensor], dtype=torch.float64):
        self.p = {k: v.to(dtype) for k, v in state_dict.items()}
        self.dtype = dtype
        self.in_channels = 4

    # ---------------------------------------------------------------- building blocks
    def _lin(self, x, name, bias=True):
        return F.linear(x, self.p[name + ".weight"], self.p.get(name + ".bias") if bias else None)

    def _conv(self, x, name, stride=1, padding=1):
        return F.conv2d(x, self.p[name + ".weight"], self.p[name + ".bias"], stride=stride, padding=padding)

    def _gn(self, x, name, eps):
        return F.group_norm(x, 32, self.p[name + ".weight"], self.p[name + ".bias"], eps)

    def _ln(self, x, name):
        return F.layer_norm(x, (x.shape[-1],), self.p[name + ".weight"], self.p[name + ".bias"], 1e-5)

    def resnet(self, x, temb, name):
        """my_diffusers/models/resnet.py:331-365 (eps 1e-5, SiLU, temb added after conv1, output_scale_factor 1)."""
        h = F.silu(self._gn(x, name + ".norm1", 1e-5))
        h = self._conv(h, name + ".conv1")
        h = h + self._lin(F.silu(temb), name + ".time_emb_proj")[:, :, None, None]
        h = F.silu(self._gn(h, name + ".norm2", 1e-5))
        h = self._conv(h, name + ".conv2")
        if name + ".conv_shortcut.weight" in self.p:
            x = self._conv(x, name + ".conv_shortcut", padding=0)
        return x + h

    def attention(self, x, ctx, name, hook, place):
        """my_diffusers/models/attention.py:250-288 with the controller call of
        models/p2p/attention_control.py:20-47 between softmax and P.V."""
        is_cross = ctx is not None
        c = ctx if is_cross else x
        q = self._lin(x, name + ".to_q", bias=False)
        k = self._lin(c, name + ".to_k", bias=False)
        v = self._lin(c, name + ".to_v", bias=False)

        def split(t):  # reshape_heads_to_batch_dim, attention.py:236-241: row index = b*8 + h
            b, n, d = t.shape
            return t.reshape(b, n, HEADS, d // HEADS).permute(0, 2, 1, 3).reshape(b * HEADS, n, d // HEADS)

        q, k, v = split(q), split(k), split(v)
        scale = (q.shape[-1]) ** -0.5
        n = q.shape[1]
        lazy = hook is None or (not is_cross and n > 32 ** 2 and hasattr(hook, "passthrough"))
        if not is_cross and n > 32 ** 2 and lazy:
            # 64x64 self-attention: the [b*8, 4096, 4096] fp64 probability tensor (1 GB per image) is never materialised.
            # Row-wise softmax makes query chunks independent, and both restated controllers are the identity on maps
            # with more than 32^2 queries (attention_control.py:223,291) - they only need their layer bookkeeping.
            if hook is not None:
                hook.passthrough(is_cross, place)
            out = torch.empty_like(q)
            for i0 in range(0, n, 512):
                sim = torch.einsum("bid,bjd->bij", q[:, i0:i0 + 512], k) * scale
                out[:, i0:i0 + 512] = torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v)
        else:
            sim = torch.einsum("bid,bjd->bij", q, k) * scale
            attn = sim.softmax(dim=-1)
            if hook is not None:
                attn = hook(attn, is_cross, place)
            out = torch.einsum("bij,bjd->bid", attn, v)
        b8, n, d = out.shape
        out = out.reshape(b8 // HEADS, HEADS, n, d).permute(0, 2, 1, 3).reshape(b8 // HEADS, n, d * HEADS)
        return self._lin(out, name + ".to_out.0")

    def transformer(self, x, ctx, name, hook, place):
        """SpatialTransformer + BasicTransformerBlock + GEGLU: attention.py:140-151,195-200,329-333."""
        b, c, h, w = x.shape
        x_in = x
        x = self._gn(x, name + ".norm", 1e-6)
        x = self._conv(x, name + ".proj_in", padding=0)
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        t = name + ".transformer_blocks.0"
        x = self.attention(self._ln(x, t + ".norm1"), None, t + ".attn1", hook, place) + x
        x = self.attention(self._ln(x, t + ".norm2"), ctx, t + ".attn2", hook, place) + x
        y = self._lin(self._ln(x, t + ".norm3"), t + ".ff.net.0.proj")
        val, gate = y.chunk(2, dim=-1)
        x = self._lin(val * F.gelu(gate), t + ".ff.net.2") + x
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return self._conv(x, name + ".proj_out", padding=0) + x_in

    # ---------------------------------------------------------------- forward
    def __call__(self, sample, timestep, encoder_hidden_states, attn_hook: Optional[Callable] = None):
        """my_diffusers/models/unet_2d_condition.py:189-273; blocks unet_blocks.py:277-367,451-612,998-1154."""
        x = sample.to(self.dtype)
        ctx = encoder_hidden_states.to(self.dtype)
        t = torch.as_tensor(timestep).reshape(-1).to(torch.float64).expand(x.shape[0])
        emb = timestep_embedding(t).to(self.dtype)
        emb = self._lin(F.silu(self._lin(emb, "time_embedding.linear_1")), "time_embedding.linear_2")
        x = self._conv(x, "conv_in")
        skips = [x]
        for i in range(4):
            for j in range(2):
                x = self.resnet(x, emb, f"down_blocks.{i}.resnets.{j}")
                if i < 3:
                    x = self.transformer(x, ctx, f"down_blocks.{i}.attentions.{j}", attn_hook, "down")
                skips.append(x)
            if i < 3:
                x = self._conv(x, f"down_blocks.{i}.downsamplers.0.conv", stride=2, padding=1)  # resnet.py:88-97
                skips.append(x)
        x = self.resnet(x, emb, "mid_block.resnets.0")
        x = self.transformer(x, ctx, "mid_block.attentions.0", attn_hook, "mid")
        x = self.resnet(x, emb, "mid_block.resnets.1")
        for i in range(4):
            for j in range(3):
                x = torch.cat([x, skips.pop()], dim=1)  # unet_blocks.py:1082,1146
                x = self.resnet(x, emb, f"up_blocks.{i}.resnets.{j}")
                if i > 0:
                    x = self.transformer(x, ctx, f"up_blocks.{i}.attentions.{j}", attn_hook, "up")
            if i < 3:
                x = F.interpolate(x, scale_factor=2.0, mode="nearest")  # resnet.py:43-50
                x = self._conv(x, f"up_blocks.{i}.upsamplers.0.conv")
        x = F.silu(self._gn(x, "conv_norm_out", 1e-5))
        return self._conv(x, "conv_out")
