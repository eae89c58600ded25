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
