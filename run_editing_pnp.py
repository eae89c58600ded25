#!/usr/bin/env python
"""Entry point with the interface of the reference's `run_editing_pnp.py` (:476-552): Plug-and-Play diffusion features
(`ddim+pnp`, `directinversion+pnp`) over a PIE-Bench mapping file, on the fused engine (pnpinversion_b200/pnp_features.py)."""
import argparse
import json

from pnpinversion_b200 import cli
from pnpinversion_b200.pnp_features import PnPFeaturesEditor


def main(argv=None):
    parser = argparse.ArgumentParser()
    cli.add_common_args(parser, ["ddim+pnp", "directinversion+pnp"])
    args = parser.parse_args(argv)
    model, _ = cli.load_model(args, max_batch=3)
    editor = PnPFeaturesEditor(model, num_ddim_steps=args.num_ddim_steps)

    def edit_one(method, item):  # run_editing_pnp.py:528-536
        return editor(method, image_path=item["image_path"], prompt_src=item["prompt_src"], prompt_tar=item["prompt_tar"],
                      guidance_scale=7.5)

    stats = cli.sweep(args, ("ddim+pnp", "directinversion+pnp"), edit_one)
    print(json.dumps({"rank": cli.dist_env()[0], **stats}))
    return stats


if __name__ == "__main__":
    main()
