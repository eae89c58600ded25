#!/usr/bin/env python
"""Entry point with the interface of the reference's `run_editing_p2p.py` (:82-146): edits every image of a PIE-Bench
`mapping_file.json` with the Prompt-to-Prompt methods of `P2PEditor` and writes the 2048x512 result strips to
<output_path>/<edit_method>/annotation_images/...; implemented on the fused B200 engine (pnpinversion_b200).

    python run_editing_p2p.py --data_path data --output_path output --edit_method_list directinversion+p2p
    torchrun --nproc-per-node 8 run_editing_p2p.py ...      # the work list is sharded over the GPUs (88/87 of 700)
"""
import argparse
import json

from pnpinversion_b200 import cli
from pnpinversion_b200.p2p_editor import SUPPORTED_METHODS, P2PEditor


def main(argv=None):
    parser = argparse.ArgumentParser()
    cli.add_common_args(parser, ["ddim+p2p"])
    args = parser.parse_args(argv)
    model, dev = cli.load_model(args, max_batch=max(4, 4 * args.batch))
    editor = P2PEditor(args.edit_method_list, dev, num_ddim_steps=args.num_ddim_steps, model=model)

    def controls(item):
        bw = item["blended_word"]
        return dict(blend_word=((bw[0],), (bw[1],)) if len(bw) else None,
                    eq_params={"words": (bw[1],), "values": (2,)} if len(bw) else None)

    def edit_one(method, item):  # the call of run_editing_p2p.py:120-138, keyword for keyword
        return editor(method, image_path=item["image_path"], prompt_src=item["prompt_src"], prompt_tar=item["prompt_tar"],
                      guidance_scale=7.5, cross_replace_steps=0.4, self_replace_steps=0.6, proximal="l0", quantile=0.75,
                      use_inversion_guidance=True, recon_lr=1, recon_t=400, **controls(item))

    def edit_many(method, items):
        if method != "directinversion+p2p":
            return [edit_one(method, it) for it in items]
        cs = [controls(it) for it in items]
        return editor.edit_batch([it["image_path"] for it in items], [it["prompt_src"] for it in items],
                                 [it["prompt_tar"] for it in items], guidance_scale=7.5, cross_replace_steps=0.4,
                                 self_replace_steps=0.6, blend_word=[c["blend_word"] for c in cs],
                                 eq_params=[c["eq_params"] for c in cs], per_image_params=True)

    stats = cli.sweep(args, SUPPORTED_METHODS + ("null-text-inversion+p2p",), edit_one, edit_many)
    print(json.dumps({"rank": cli.dist_env()[0], **stats}))
    return stats


if __name__ == "__main__":
    main()
