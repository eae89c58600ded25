"""Shared driver of the `run_editing_*` entry points: the PIE-Bench `mapping_file.json` sweep of the reference's scripts
(`run_editing_p2p.py:82-146`, `run_editing_masactrl.py:179-234`, `run_editing_edict.py:63-120`) with the same flags,
dataset schema, output layout and skip-if-exists behaviour -- plus what the reference leaves to hand-started copies of
the script: image-parallel sharding of the work list over the GPUs of one box (`torchrun --nproc-per-node N`, contiguous
uneven split through parallel.shard_bounds: 700 images on 8 GPUs -> 88,88,88,88,87,87,87,87).

On-disk formats (README.md:131-143 of the reference):
  <data_path>/mapping_file.json   {key: {"image_path", "original_prompt", "editing_prompt", "editing_instruction",
                                         "editing_type_id", "blended_word", "mask" (RLE: [start, length, ...])}}
  <data_path>/annotation_images/<image_path>
  <output_path>/<edit_method>/annotation_images/<image_path>   the 2048x512 strip [instruction|source|reconstruction|edit]
"""
from __future__ import annotations

import argparse
import json
import os
import random
from typing import Callable, Dict, Iterable, List, Sequence

import numpy as np

CATEGORIES = [str(i) for i in range(10)]


def mask_decode(encoded_mask: Sequence[int], image_shape=(512, 512)) -> np.ndarray:
    """run_editing_p2p.py:11-27: run-length pairs (start, length) over the flattened image -> {0,1} mask whose border
    is forced to 1 ("to avoid annotation errors in boundary")."""
    length = image_shape[0] * image_shape[1]
    flat = np.zeros((length,))
    for start, run in zip(encoded_mask[0::2], encoded_mask[1::2]):
        flat[start:start + min(run, length - start)] = 1
    m = flat.reshape(image_shape[0], image_shape[1])
    m[0, :] = m[-1, :] = 1
    m[:, 0] = m[:, -1] = 1
    return m


def mask_encode(mask: np.ndarray) -> List[int]:
    """Inverse of mask_decode's run-length format (used by the synthetic dataset writer and the tests)."""
    flat = np.asarray(mask).reshape(-1).astype(np.int8)
    d = np.diff(np.concatenate([[0], flat, [0]]))
    starts, ends = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
    out: List[int] = []
    for s, e in zip(starts, ends):
        out += [int(s), int(e - s)]
    return out


def setup_seed(seed: int = 1234) -> None:
    """run_editing_p2p.py:30-36."""
    import torch

    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def add_common_args(parser: argparse.ArgumentParser, default_methods: List[str]) -> None:
    parser.add_argument("--rerun_exist_images", action="store_true")
    parser.add_argument("--data_path", type=str, default="data")
    parser.add_argument("--output_path", type=str, default="output")
    parser.add_argument("--edit_category_list", nargs="+", type=str, default=list(CATEGORIES))
    parser.add_argument("--edit_method_list", nargs="+", type=str, default=default_methods)
    # additions of this implementation
    parser.add_argument("--checkpoint", type=str, default=os.environ.get("PNP_SD_CHECKPOINT"),
                        help="local SD-1.x diffusers directory (unet/ vae/ text_encoder/ tokenizer/); without it the "
                             "seeded random-init stand-in is used (no network here)")
    parser.add_argument("--num_ddim_steps", type=int, default=50)
    parser.add_argument("--batch", type=int, default=1, help="images that share every UNet call (directinversion methods)")
    parser.add_argument("--limit", type=int, default=0, help="edit at most this many work items (0 = all)")


def read_items(data_path: str, categories: Iterable[str]) -> List[Dict]:
    """The work list in mapping-file order (run_editing_p2p.py:99-112), brackets stripped from the prompts."""
    with open(os.path.join(data_path, "mapping_file.json"), "r") as f:
        instr = json.load(f)
    cats = set(categories)
    items = []
    for key, item in instr.items():
        if item["editing_type_id"] not in cats:
            continue
        blended = item["blended_word"].split(" ") if item["blended_word"] != "" else []
        items.append(dict(key=key, prompt_src=item["original_prompt"].replace("[", "").replace("]", ""),
                          prompt_tar=item["editing_prompt"].replace("[", "").replace("]", ""),
                          image_path=os.path.join(f"{data_path}/annotation_images", item["image_path"]),
                          editing_instruction=item["editing_instruction"], blended_word=blended,
                          mask=item.get("mask", [])))
    return items


def out_path(item: Dict, data_path: str, output_path: str, method: str) -> str:
    """run_editing_p2p.py:115: the data prefix of the image path is replaced by <output_path>/<method>."""
    return item["image_path"].replace(data_path, os.path.join(output_path, method))


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def my_shard(items: List, rank: int, world: int) -> List:
    from .parallel import shard_bounds

    lo, hi = shard_bounds(len(items), rank, world)
    return items[lo:hi]


def sweep(args, methods_known: Sequence[str], edit_one: Callable, edit_many: Callable = None) -> Dict[str, int]:
    """The reference's double loop (items x methods) over THIS rank's shard.  `edit_one(method, item) -> PIL image`;
    `edit_many(method, [items]) -> [PIL images]` serves --batch > 1 where a method supports it."""
    import torch

    rank, world, _ = dist_env()
    for m in args.edit_method_list:
        if m not in methods_known:
            raise NotImplementedError(f"No edit method named {m}")
    items = read_items(args.data_path, args.edit_category_list)
    if args.limit:
        items = items[: args.limit]
    mine = my_shard(items, rank, world)
    stats = {"edited": 0, "skipped": 0, "items": len(mine), "total_items": len(items)}
    for method in args.edit_method_list:
        todo = []
        for it in mine:
            dst = out_path(it, args.data_path, args.output_path, method)
            if os.path.exists(dst) and not args.rerun_exist_images:
                print(f"skip image [{it['image_path']}] with [{method}]")
                stats["skipped"] += 1
            else:
                todo.append((it, dst))
        step = max(1, args.batch) if edit_many is not None else 1
        for i in range(0, len(todo), step):
            chunk = todo[i:i + step]
            for it, _ in chunk:
                print(f"editing image [{it['image_path']}] with [{method}]")
            setup_seed()
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
            if len(chunk) > 1:
                images = edit_many(method, [it for it, _ in chunk])
            else:
                images = [edit_one(method, chunk[0][0])]
            for (it, dst), im in zip(chunk, images):
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                im.save(dst)
                stats["edited"] += 1
            print("finish")
    return stats


def load_model(args, max_batch: int = 4):
    """StableDiffusionPipeline.from_pretrained(...) of the reference editors: a local checkpoint directory, or (offline)
    the seeded random-init UNet + VAE with the whitespace tokenizer."""
    import torch

    from .model import FusedModel

    _, _, local = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("the run_editing_* entry points need a CUDA device (sm_100a); there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if args.checkpoint:
        return FusedModel.from_pretrained(args.checkpoint, device=dev, max_batch=max_batch), dev
    print("[pnpinversion_b200] no --checkpoint / PNP_SD_CHECKPOINT: using the seeded random-init SD-1.x stand-in")
    return FusedModel.synthetic(device=dev, max_batch=max_batch, with_vae=True), dev


def write_synthetic_dataset(root: str, n_items: int = 8, size: int = 512, seed: int = 0) -> str:
    """A PIE-Bench-shaped dataset (mapping_file.json + annotation_images/<category>/<id>.jpg) of seeded synthetic images,
    for exercising the entry points offline; categories cycle through 0..9."""
    from PIL import Image

    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "annotation_images"), exist_ok=True)
    mapping = {}
    yy, xx = np.mgrid[0:size, 0:size]
    for i in range(n_items):
        cat = str(i % 10)
        rel = os.path.join(f"{cat}_synthetic", f"{i:012d}.jpg")
        os.makedirs(os.path.join(root, "annotation_images", os.path.dirname(rel)), exist_ok=True)
        q = max(size // 4, 2)
        cx, cy, r = rng.randint(q, size - q), rng.randint(q, size - q), rng.randint(max(q // 3, 1), q)
        base = np.stack([(yy * (i + 1)) % 256, (xx * 2) % 256, (yy + xx) % 256], axis=-1).astype(np.float32)
        disk = ((yy - cy) ** 2 + (xx - cx) ** 2) < r * r
        base[disk] = rng.randint(0, 255, 3)
        Image.fromarray(base.clip(0, 255).astype(np.uint8)).save(os.path.join(root, "annotation_images", rel), quality=95)
        mapping[f"{i:012d}"] = {
            "image_path": rel, "original_prompt": "a [cat] sitting on a table with a green eyes",
            "editing_prompt": "a [dog] sitting on a table with a green eyes", "editing_instruction": "change the cat to a dog",
            "editing_type_id": cat, "blended_word": "cat dog", "mask": mask_encode(disk.astype(np.uint8)),
        }
    with open(os.path.join(root, "mapping_file.json"), "w") as f:
        json.dump(mapping, f)
    return root
