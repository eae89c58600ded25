"""SD-1.x UNet2DConditionModel architecture tables (host side).

The arithmetic spec is the reference's vendored diffusers 0.3.0
(`models/edict/my_diffusers/models/unet_2d_condition.py:58-165`, `unet_blocks.py:277-367,451-612,998-1154`);
the key names are those of an SD-1.x diffusers checkpoint.  The same table is
compiled into the C++ engine (`csrc/unet_engine.cu`); `tests/test_capi_cpu.py` checks the two agree.
"""
from __future__ import annotations

from typing import List, Tuple

BLOCK_OUT = (320, 640, 1280, 1280)
LAYERS_PER_BLOCK = 2
CROSS_DIM = 768
TIME_DIM = 1280
HEADS = 8
GROUPS = 32
LATENT_C = 4
MAX_TOKENS = 77

Spec = Tuple[str, Tuple[int, ...]]


def _resnet(specs: List[Spec], p: str, cin: int, cout: int) -> None:
    specs += [
        (f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)),
        (f"{p}.conv1.weight", (cout, cin, 3, 3)), (f"{p}.conv1.bias", (cout,)),
        (f"{p}.time_emb_proj.weight", (cout, TIME_DIM)), (f"{p}.time_emb_proj.bias", (cout,)),
        (f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)),
        (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,)),
    ]
    if cin != cout:
        specs += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]


def _transformer(specs: List[Spec], p: str, c: int) -> None:
    t = f"{p}.transformer_blocks.0"
    specs += [
        (f"{p}.norm.weight", (c,)), (f"{p}.norm.bias", (c,)),
        (f"{p}.proj_in.weight", (c, c, 1, 1)), (f"{p}.proj_in.bias", (c,)),
        (f"{t}.attn1.to_q.weight", (c, c)), (f"{t}.attn1.to_k.weight", (c, c)), (f"{t}.attn1.to_v.weight", (c, c)),
        (f"{t}.attn1.to_out.0.weight", (c, c)), (f"{t}.attn1.to_out.0.bias", (c,)),
        (f"{t}.ff.net.0.proj.weight", (8 * c, c)), (f"{t}.ff.net.0.proj.bias", (8 * c,)),
        (f"{t}.ff.net.2.weight", (c, 4 * c)), (f"{t}.ff.net.2.bias", (c,)),
        (f"{t}.attn2.to_q.weight", (c, c)), (f"{t}.attn2.to_k.weight", (c, CROSS_DIM)),
        (f"{t}.attn2.to_v.weight", (c, CROSS_DIM)),
        (f"{t}.attn2.to_out.0.weight", (c, c)), (f"{t}.attn2.to_out.0.bias", (c,)),
        (f"{t}.norm1.weight", (c,)), (f"{t}.norm1.bias", (c,)),
        (f"{t}.norm2.weight", (c,)), (f"{t}.norm2.bias", (c,)),
        (f"{t}.norm3.weight", (c,)), (f"{t}.norm3.bias", (c,)),
        (f"{p}.proj_out.weight", (c, c, 1, 1)), (f"{p}.proj_out.bias", (c,)),
    ]


def unet_param_specs() -> List[Spec]:
    """All 686 parameter tensors of the SD-1.x UNet, (name, shape)."""
    s: List[Spec] = [
        ("conv_in.weight", (BLOCK_OUT[0], LATENT_C, 3, 3)), ("conv_in.bias", (BLOCK_OUT[0],)),
        ("time_embedding.linear_1.weight", (TIME_DIM, BLOCK_OUT[0])), ("time_embedding.linear_1.bias", (TIME_DIM,)),
        ("time_embedding.linear_2.weight", (TIME_DIM, TIME_DIM)), ("time_embedding.linear_2.bias", (TIME_DIM,)),
    ]
    # down blocks: CrossAttnDown x3, DownBlock2D
    cin = BLOCK_OUT[0]
    for i, cout in enumerate(BLOCK_OUT):
        has_attn = i < 3
        for j in range(LAYERS_PER_BLOCK):
            _resnet(s, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
            if has_attn:
                _transformer(s, f"down_blocks.{i}.attentions.{j}", cout)
        if i < 3:
            s += [(f"down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"down_blocks.{i}.downsamplers.0.conv.bias", (cout,))]
        cin = cout
    # mid
    c = BLOCK_OUT[-1]
    _resnet(s, "mid_block.resnets.0", c, c)
    _transformer(s, "mid_block.attentions.0", c)
    _resnet(s, "mid_block.resnets.1", c, c)
    # up blocks: UpBlock2D, CrossAttnUp x3
    rev = list(reversed(BLOCK_OUT))
    prev = rev[0]
    for i, cout in enumerate(rev):
        cin_blk = rev[min(i + 1, 3)]
        has_attn = i > 0
        for j in range(LAYERS_PER_BLOCK + 1):
            skip = cin_blk if j == LAYERS_PER_BLOCK else cout
            rin = prev if j == 0 else cout
            _resnet(s, f"up_blocks.{i}.resnets.{j}", rin + skip, cout)
            if has_attn:
                _transformer(s, f"up_blocks.{i}.attentions.{j}", cout)
        if i < 3:
            s += [(f"up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
        prev = cout
    s += [("conv_norm_out.weight", (BLOCK_OUT[0],)), ("conv_norm_out.bias", (BLOCK_OUT[0],)),
          ("conv_out.weight", (LATENT_C, BLOCK_OUT[0], 3, 3)), ("conv_out.bias", (LATENT_C,))]
    return s


# Execution order of the 16 transformer blocks (= 32 attention layers, self then cross), with the
# `place_in_unet` the reference's registration assigns (models/p2p/attention_control.py:71-81) and the token count.
def transformer_order(latent_hw: int = 64):
    order = []
    hw = latent_hw
    for i in range(3):
        for j in range(2):
            order.append((f"down_blocks.{i}.attentions.{j}", "down", BLOCK_OUT[i], hw * hw))
        hw //= 2
    order.append(("mid_block.attentions.0", "mid", BLOCK_OUT[3], (latent_hw // 8) ** 2))
    hw = latent_hw // 4
    for i in range(1, 4):
        for j in range(3):
            order.append((f"up_blocks.{i}.attentions.{j}", "up", list(reversed(BLOCK_OUT))[i], hw * hw))
        hw *= 2
    return order


# ---------------------------------------------------------------------------------------------------- AutoencoderKL
# SD-1.x VAE ("next"): `models/edict/my_diffusers/models/vae.py:54-131,133-210,480-557`,
# blocks `unet_blocks.py` (DownEncoderBlock2D / UpDecoderBlock2D / UNetMidBlock2D), `attention.py:9-93` (AttentionBlock).
VAE_BLOCK_OUT = (128, 256, 512, 512)
VAE_LAYERS_PER_BLOCK = 2
VAE_LATENT_C = 4


def _vae_resnet(specs: List[Spec], p: str, cin: int, cout: int) -> None:
    specs += [
        (f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)),
        (f"{p}.conv1.weight", (cout, cin, 3, 3)), (f"{p}.conv1.bias", (cout,)),
        (f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)),
        (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,)),
    ]
    if cin != cout:
        specs += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]


def _vae_mid(specs: List[Spec], p: str, c: int) -> None:
    a = f"{p}.attentions.0"
    specs += [(f"{a}.group_norm.weight", (c,)), (f"{a}.group_norm.bias", (c,))]
    for n in ("query", "key", "value", "proj_attn"):
        specs += [(f"{a}.{n}.weight", (c, c)), (f"{a}.{n}.bias", (c,))]
    _vae_resnet(specs, f"{p}.resnets.0", c, c)
    _vae_resnet(specs, f"{p}.resnets.1", c, c)


def vae_param_specs() -> List[Spec]:
    """(name, shape) of the 248 AutoencoderKL parameters (83.65 M values); names as in a diffusers SD-1.x `vae/` state dict."""
    bo = VAE_BLOCK_OUT
    s: List[Spec] = []
    # encoder (vae.py:54-131)
    s += [("encoder.conv_in.weight", (bo[0], 3, 3, 3)), ("encoder.conv_in.bias", (bo[0],))]
    cout = bo[0]
    for i in range(len(bo)):
        cin, cout = cout, bo[i]
        for j in range(VAE_LAYERS_PER_BLOCK):
            _vae_resnet(s, f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i != len(bo) - 1:
            s += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (cout,))]
    _vae_mid(s, "encoder.mid_block", bo[-1])
    s += [("encoder.conv_norm_out.weight", (bo[-1],)), ("encoder.conv_norm_out.bias", (bo[-1],)),
          ("encoder.conv_out.weight", (2 * VAE_LATENT_C, bo[-1], 3, 3)), ("encoder.conv_out.bias", (2 * VAE_LATENT_C,))]
    # decoder (vae.py:133-210)
    s += [("decoder.conv_in.weight", (bo[-1], VAE_LATENT_C, 3, 3)), ("decoder.conv_in.bias", (bo[-1],))]
    _vae_mid(s, "decoder.mid_block", bo[-1])
    rev = list(reversed(bo))
    cout = rev[0]
    for i in range(len(rev)):
        cin, cout = cout, rev[i]
        for j in range(VAE_LAYERS_PER_BLOCK + 1):
            _vae_resnet(s, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i != len(rev) - 1:
            s += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
    s += [("decoder.conv_norm_out.weight", (bo[0],)), ("decoder.conv_norm_out.bias", (bo[0],)),
          ("decoder.conv_out.weight", (3, bo[0], 3, 3)), ("decoder.conv_out.bias", (3,))]
    s += [("quant_conv.weight", (2 * VAE_LATENT_C, 2 * VAE_LATENT_C, 1, 1)), ("quant_conv.bias", (2 * VAE_LATENT_C,)),
          ("post_quant_conv.weight", (VAE_LATENT_C, VAE_LATENT_C, 1, 1)), ("post_quant_conv.bias", (VAE_LATENT_C,))]
    return s
