"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the SD-1.x AutoencoderKL the reference calls around the editing loop
(`utils/utils.py:58-80`: `image2latent` = encode(img).latent_dist.mean * 0.18215, `latent2image` = decode(z / 0.18215);
one encode and five decodes per edited image).

Arithmetic spec: the reference's vendored diffusers 0.3.0, `models/edict/my_diffusers/models/vae.py:54-131` (Encoder),
`:133-210` (Decoder), `:480-557` (AutoencoderKL: quant_conv / post_quant_conv), `unet_blocks.py` DownEncoderBlock2D /
UpDecoderBlock2D / UNetMidBlock2D (resnets without time embedding, GroupNorm eps 1e-6), `resnet.py:64-97` (Downsample2D with
padding 0 = pad right/bottom by one, stride 2), `:16-52` (nearest 2x + conv), `attention.py:9-93` (single-head
AttentionBlock, softmax in float64).

PARITY PINNED: tests/golden/vae_small.npz holds outputs of the vendored AutoencoderKL itself (oracle/make_golden.py vae) on
the synthetic weights of pnpinversion_b200/synth.py; tests/test_oracle_cpu.py checks this file against it.  The CUDA VAE is
not built yet - this oracle is the gate it will have to pass.  Only tests/ may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from pnpinversion_b200.arch import VAE_BLOCK_OUT, VAE_LAYERS_PER_BLOCK


class VaeRef:
    def __init__(self, state_dict, dtype=torch.float64):
        self.p = {k: v.to(dtype) for k, v in state_dict.items()}
        self.dtype = dtype

    # ---- building blocks
    def _conv(self, x, name, stride=1, padding=1):
        return F.conv2d(x, self.p[name + ".weight"], self.p[name + ".bias"], stride=stride, padding=padding)

    def _gn(self, x, name):
        return F.group_norm(x, 32, self.p[name + ".weight"], self.p[name + ".bias"], eps=1e-6)

    def _resnet(self, x, name):
        h = self._conv(F.silu(self._gn(x, name + ".norm1")), name + ".conv1")
        h = self._conv(F.silu(self._gn(h, name + ".norm2")), name + ".conv2")
        if name + ".conv_shortcut.weight" in self.p:
            x = self._conv(x, name + ".conv_shortcut", padding=0)
        return x + h

    def _attn(self, x, name):
        b, c, hh, ww = x.shape
        h = self._gn(x, name + ".group_norm").reshape(b, c, hh * ww).transpose(1, 2)
        q = F.linear(h, self.p[name + ".query.weight"], self.p[name + ".query.bias"])
        k = F.linear(h, self.p[name + ".key.weight"], self.p[name + ".key.bias"])
        v = F.linear(h, self.p[name + ".value.weight"], self.p[name + ".value.bias"])
        scale = c ** -0.25  # one head: q and k are each scaled by 1 / sqrt(sqrt(c))
        probs = torch.softmax(((q * scale) @ (k * scale).transpose(1, 2)).double(), dim=-1).to(h.dtype)
        o = F.linear(probs @ v, self.p[name + ".proj_attn.weight"], self.p[name + ".proj_attn.bias"])
        return o.transpose(1, 2).reshape(b, c, hh, ww) + x

    def _mid(self, x, name):
        x = self._resnet(x, name + ".resnets.0")
        x = self._attn(x, name + ".attentions.0")
        return self._resnet(x, name + ".resnets.1")

    # ---- the two entry points the reference uses
    def encode_moments(self, img):
        """(mean, logvar) of the posterior, each (B,4,H/8,W/8); `image2latent` takes the mean."""
        x = self._conv(img.to(self.dtype), "encoder.conv_in")
        n = len(VAE_BLOCK_OUT)
        for i in range(n):
            for j in range(VAE_LAYERS_PER_BLOCK):
                x = self._resnet(x, f"encoder.down_blocks.{i}.resnets.{j}")
            if i != n - 1:
                x = self._conv(F.pad(x, (0, 1, 0, 1)), f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=2, padding=0)
        x = self._mid(x, "encoder.mid_block")
        x = self._conv(F.silu(self._gn(x, "encoder.conv_norm_out")), "encoder.conv_out")
        m = self._conv(x, "quant_conv", padding=0)
        return m[:, :4], m[:, 4:]

    def decode(self, z):
        x = self._conv(z.to(self.dtype), "post_quant_conv", padding=0)
        x = self._conv(x, "decoder.conv_in")
        x = self._mid(x, "decoder.mid_block")
        n = len(VAE_BLOCK_OUT)
        for i in range(n):
            for j in range(VAE_LAYERS_PER_BLOCK + 1):
                x = self._resnet(x, f"decoder.up_blocks.{i}.resnets.{j}")
            if i != n - 1:
                x = self._conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), f"decoder.up_blocks.{i}.upsamplers.0.conv")
        return self._conv(F.silu(self._gn(x, "decoder.conv_norm_out")), "decoder.conv_out")
