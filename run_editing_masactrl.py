#!/usr/bin/env python
"""Entry point with the interface of the reference's `run_editing_masactrl.py` (:179-234): MasaCtrl mutual
self-attention editing (`ddim+masactrl`, `directinversion+masactrl`) over a PIE-Bench mapping file, on the fused engine."""
import argparse
import json

from pnpinversion_b200 import cli
from pnpinversion_b200.masactrl import MasaCtrlEditor


def main(argv=None):
    parser = argparse.ArgumentParser()
    cli.add_common_args(parser, ["ddim+masactrl", "directinversion+masactrl"])
    args = parser.parse_args(argv)
    model, dev = cli.load_model(args, max_batch=max(4, 4 * args.batch))
    editor = MasaCtrlEditor(args.edit_method_list, dev, num_ddim_steps=args.num_ddim_steps, model=model)

    def edit_one(method, item):  # run_editing_masactrl.py:218-225
        return editor(method, image_path=item["image_path"], prompt_src=item["prompt_src"], prompt_tar=item["prompt_tar"],
                      guidance_scale=7.5, step=4, layper=10)

    def edit_many(method, items):
        if method != "directinversion+masactrl":
            return [edit_one(method, it) for it in items]
        return editor.edit_batch_images([it["image_path"] for it in items], [it["prompt_src"] for it in items],
                                        [it["prompt_tar"] for it in items], guidance_scale=7.5, step=4, layper=10)

    stats = cli.sweep(args, ("ddim+masactrl", "directinversion+masactrl"), edit_one, edit_many)
    print(json.dumps({"rank": cli.dist_env()[0], **stats}))
    return stats


if __name__ == "__main__":
    main()
