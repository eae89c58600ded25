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
    save_file({"model.diffusion_model.input_blocks.0.0.weight": torch.zeros(1)}, f)
    with pytest.raises(NotImplementedError, match="LDM"):
        checkpoint.load_unet_state_dict(f)


def test_text_components_are_optional(tmp_path):
    assert checkpoint.load_text_components(str(tmp_path)) == (None, None)
