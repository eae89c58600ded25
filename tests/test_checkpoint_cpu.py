"""Checkpoint reading / validation (pnpinversion_b200/checkpoint.py): host logic, no GPU."""
import os

import pytest
import torch

from pnpinversion_b200 import arch, checkpoint


def _tiny_sd(n=6):
    # the first few tensors of the table with their real shapes (a whole UNet is 1.7 GB: too big for a unit test)
    return {k: torch.full(shape, float(i), dtype=torch.float16) for i, (k, shape) in enumerate(arch.unet_param_specs()[:n])}


def test_find_and_read_safetensors_and_bin_in_the_diffusers_layout(tmp_path):
    from safetensors.torch import save_file

    sd = _tiny_sd()
    pipe = tmp_path / "sd-v1-4"
    (pipe / "unet").mkdir(parents=True)
    save_file(sd, str(pipe / "unet" / "diffusion_pytorch_model.safetensors"))
    torch.save(sd, str(pipe / "unet" / "diffusion_pytorch_model.bin"))
    f = checkpoint.find_unet_file(str(pipe))  # pipeline directory -> unet/, safetensors preferred
    assert f.endswith(os.path.join("unet", "diffusion_pytorch_model.safetensors"))
    assert checkpoint.find_unet_file(str(pipe / "unet")) == f and checkpoint.find_unet_file(f) == f
    for path in (f, str(pipe / "unet" / "diffusion_pytorch_model.bin")):
        got = checkpoint.read_state_dict(path)
        assert sorted(got) == sorted(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        checkpoint.find_unet_file(str(tmp_path / "nothing-here"))


def test_validation_reports_missing_unexpected_and_misshaped_tensors(tmp_path):
    specs = arch.unet_param_specs()
    assert len(specs) == 686
    # shape-only stand-ins (meta tensors) for a complete state dict: validation never touches the data
    full = {k: torch.empty(shape, device="meta") for k, shape in specs}
    assert checkpoint.check_unet_state_dict(full) == ([], [], [])
    broken = dict(full)
    k0, k1 = specs[0][0], specs[10][0]
    del broken[k0]
    broken[k1] = torch.empty((3, 3), device="meta")
    broken["some.extra.weight"] = torch.empty((1,), device="meta")
    missing, unexpected, bad = checkpoint.check_unet_state_dict(broken)
    assert missing == [k0] and unexpected == ["some.extra.weight"] and len(bad) == 1 and bad[0].startswith(k1 + ":")

    from safetensors.torch import save_file

    f = str(tmp_path / "diffusion_pytorch_model.safetensors")
    save_file(_tiny_sd(), f)
    with pytest.raises(ValueError, match=r"680 missing"):
        checkpoint.load_unet_state_dict(f)


def test_text_components_are_optional(tmp_path):
    assert checkpoint.load_text_components(str(tmp_path)) == (None, None)


def _to_ldm(name):
    """diffusers -> original CompVis / LDM key (the inverse of checkpoint.ldm_to_diffusers_unet), SD-1.x layout."""
    inv = {v: k for k, v in checkpoint._LDM_RESNET.items()}
    p = name.split(".")

    def res(prefix, rest):
        head, _, tail = rest.partition(".")
        return f"{prefix}.{inv[head]}.{tail}"

    if p[0] == "time_embedding":
        return f"time_embed.{0 if p[1] == 'linear_1' else 2}.{p[2]}"
    if p[0] in ("conv_in", "conv_norm_out", "conv_out"):
        return {"conv_in": "input_blocks.0.0", "conv_norm_out": "out.0", "conv_out": "out.2"}[p[0]] + "." + p[1]
    if p[0] == "down_blocks":
        b = int(p[1])
        if p[2] == "resnets":
            return res(f"input_blocks.{1 + 3 * b + int(p[3])}.0", ".".join(p[4:]))
        if p[2] == "attentions":
            return f"input_blocks.{1 + 3 * b + int(p[3])}.1." + ".".join(p[4:])
        return f"input_blocks.{3 * (b + 1)}.0.op." + ".".join(p[5:])
    if p[0] == "mid_block":
        if p[1] == "attentions":
            return "middle_block.1." + ".".join(p[3:])
        return res(f"middle_block.{0 if p[2] == '0' else 2}", ".".join(p[3:]))
    b = int(p[1])
    if p[2] == "resnets":
        return res(f"output_blocks.{3 * b + int(p[3])}.0", ".".join(p[4:]))
    if p[2] == "attentions":
        return f"output_blocks.{3 * b + int(p[3])}.1." + ".".join(p[4:])
    return f"output_blocks.{3 * b + 2}.{1 if b == 0 else 2}." + ".".join(p[4:])


def test_original_ldm_single_file_layout_is_converted(tmp_path):
    """`model.diffusion_model.*` keys of a CompVis checkpoint map onto exactly the 686 diffusers names with the right shapes;
    VAE / CLIP / other tensors of the single file are ignored.  Spot checks of well-known key pairs guard the inverse
    mapping used to build the input."""
    specs = arch.unet_param_specs()
    assert _to_ldm("down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight") == \
        "input_blocks.2.1.transformer_blocks.0.attn2.to_k.weight"
    assert _to_ldm("down_blocks.1.downsamplers.0.conv.weight") == "input_blocks.6.0.op.weight"
    assert _to_ldm("up_blocks.0.upsamplers.0.conv.bias") == "output_blocks.2.1.conv.bias"
    assert _to_ldm("up_blocks.2.upsamplers.0.conv.bias") == "output_blocks.8.2.conv.bias"
    assert _to_ldm("up_blocks.3.resnets.2.conv_shortcut.weight") == "output_blocks.11.0.skip_connection.weight"
    assert _to_ldm("mid_block.resnets.1.time_emb_proj.bias") == "middle_block.2.emb_layers.1.bias"
    assert _to_ldm("time_embedding.linear_2.weight") == "time_embed.2.weight"
    ldm = {"model.diffusion_model." + _to_ldm(k): torch.empty(shape, device="meta") for k, shape in specs}
    assert len(ldm) == 686  # the inverse mapping is injective
    ldm["first_stage_model.decoder.conv_in.weight"] = torch.empty((1,), device="meta")
    ldm["cond_stage_model.transformer.text_model.embeddings.position_ids"] = torch.empty((1, 77), device="meta")
    back = checkpoint.ldm_to_diffusers_unet(ldm)
    assert checkpoint.check_unet_state_dict(back) == ([], [], [])
    with pytest.raises(ValueError):
        checkpoint.ldm_to_diffusers_unet({"model.diffusion_model.label_emb.0.0.weight": torch.empty((1,), device="meta")})


def test_clip_text_encoder_weights_are_read_and_validated(tmp_path):
    """`<checkpoint>/text_encoder/model.safetensors` -> the tensors the fused CLIP text encoder loads (clip.py), whatever
    the layer count / vocabulary of the file; `position_ids` buffers of older transformers versions are ignored, a missing
    or mis-shaped tensor is reported by name.  Also written by a live `transformers.CLIPTextModel` when importable."""
    from safetensors.torch import save_file

    from pnpinversion_b200 import synth
    from pnpinversion_b200.clip import clip_text_param_specs

    enc = tmp_path / "text_encoder"
    enc.mkdir()
    sd = synth.synth_clip_state_dict(0, layers=2, vocab=64)
    full = dict(sd)
    full["text_model.embeddings.position_ids"] = torch.arange(77).unsqueeze(0)
    save_file({k: v.contiguous() for k, v in full.items()}, str(enc / "model.safetensors"))
    got = checkpoint.load_clip_state_dict(str(enc))
    assert list(got) == [n for n, _ in clip_text_param_specs(2, 64)]
    assert all(torch.equal(got[k], sd[k]) for k in got)
    bad = dict(sd)
    del bad["text_model.encoder.layers.1.mlp.fc2.bias"]
    save_file({k: v.contiguous() for k, v in bad.items()}, str(enc / "model.safetensors"))
    with pytest.raises(ValueError, match="layers.1.mlp.fc2.bias missing"):
        checkpoint.load_clip_state_dict(str(enc))
    bad = dict(sd)
    bad["text_model.final_layer_norm.weight"] = torch.zeros(10)
    save_file({k: v.contiguous() for k, v in bad.items()}, str(enc / "model.safetensors"))
    with pytest.raises(ValueError, match="final_layer_norm.weight has shape"):
        checkpoint.load_clip_state_dict(str(enc))
    with pytest.raises(FileNotFoundError):
        checkpoint.load_clip_state_dict(str(tmp_path / "text_encoder_absent"))
    try:
        from transformers import CLIPTextConfig, CLIPTextModel
    except Exception:
        return
    cfg = CLIPTextConfig(vocab_size=64, hidden_size=768, intermediate_size=3072, num_hidden_layers=2,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu")
    CLIPTextModel(cfg).save_pretrained(str(enc))
    live = checkpoint.load_clip_state_dict(str(enc))
    assert {k: tuple(v.shape) for k, v in live.items()} == dict(clip_text_param_specs(2, 64))
