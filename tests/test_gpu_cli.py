"""The run_editing_* entry points end to end on a GPU: a PIE-Bench-shaped synthetic dataset (mapping_file.json +
annotation_images/) goes in, the reference's output tree of 2048x512 strips comes out (run_editing_p2p.py:115-141)."""
import os

import pytest
from PIL import Image

from pnpinversion_b200 import cli

pytestmark = pytest.mark.gpu


def _check_tree(items, data, out, method):
    for it in items:
        dst = cli.out_path(it, data, out, method)
        assert os.path.exists(dst), dst
        assert Image.open(dst).size == (2048, 512)


def test_run_editing_p2p_cli(cuda, tmp_path):
    import run_editing_p2p

    data, out = str(tmp_path / "data"), str(tmp_path / "output")
    cli.write_synthetic_dataset(data, n_items=3, size=512)
    argv = ["--data_path", data, "--output_path", out, "--num_ddim_steps", "3", "--batch", "2", "--edit_method_list",
            "directinversion+p2p", "ddim+p2p"]
    st = run_editing_p2p.main(argv)
    assert st["edited"] == 6 and st["skipped"] == 0
    items = cli.read_items(data, cli.CATEGORIES)
    _check_tree(items, data, out, "directinversion+p2p")
    _check_tree(items, data, out, "ddim+p2p")
    st = run_editing_p2p.main(argv)  # second run: everything exists
    assert st["edited"] == 0 and st["skipped"] == 6
    with pytest.raises(NotImplementedError):
        run_editing_p2p.main(["--data_path", data, "--output_path", out, "--edit_method_list", "no-such-method"])


def test_run_editing_masactrl_and_edict_cli(cuda, tmp_path):
    import run_editing_edict
    import run_editing_masactrl

    data, out = str(tmp_path / "data"), str(tmp_path / "output")
    cli.write_synthetic_dataset(data, n_items=2, size=512)
    items = cli.read_items(data, cli.CATEGORIES)
    st = run_editing_masactrl.main(["--data_path", data, "--output_path", out, "--num_ddim_steps", "6", "--batch", "2",
                                    "--edit_method_list", "directinversion+masactrl", "ddim+masactrl"])
    assert st["edited"] == 4
    _check_tree(items, data, out, "directinversion+masactrl")
    _check_tree(items, data, out, "ddim+masactrl")
    st = run_editing_edict.main(["--data_path", data, "--output_path", out, "--num_ddim_steps", "5", "--limit", "1",
                                 "--edit_method_list", "edict+p2p", "edict+direct_forward"])
    assert st["edited"] == 2
    _check_tree(items[:1], data, out, "edict+p2p")
