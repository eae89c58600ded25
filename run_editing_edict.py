#!/usr/bin/env python
"""Entry point with the interface of the reference's `run_editing_edict.py` (:63-120): EDICT coupled exact inversion
(`edict+direct_forward`, `edict+p2p`) over a PIE-Bench mapping file, on the fused engine."""
import argparse
import json

from pnpinversion_b200 import cli
from pnpinversion_b200.edict import edit_image_edict_p2p_strip


def main(argv=None):
    parser = argparse.ArgumentParser()
    cli.add_common_args(parser, ["edict+p2p"])
    args = parser.parse_args(argv)
    model, _ = cli.load_model(args, max_batch=max(4, 3 * args.batch))

    def edit_one(method, item):  # run_editing_edict.py:100-112
        return edit_image_edict_p2p_strip(model, item["image_path"], item["prompt_src"], item["prompt_tar"],
                                          use_p2p=(method == "edict+p2p"), steps=args.num_ddim_steps)

    stats = cli.sweep(args, ("edict+direct_forward", "edict+p2p"), edit_one)
    print(json.dumps({"rank": cli.dist_env()[0], **stats}))
    return stats


if __name__ == "__main__":
    main()
