"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference implementation.

This module imports cure-lab/PnPInversion's own Python (``/root/reference``) so that the
restated oracle (``oracle/unet_ref.py``, ``oracle/p2p_ref.py``) and the product can be pinned
against outputs of the reference itself.  It only works in the build container (the GPU box has
no ``/root/reference``); everything it produces is committed as fixtures under ``tests/golden/``
by ``oracle/make_golden.py``.  Nothing in the product imports this file.

How the reference is made importable without modifying it:
  * ``models/edict/my_diffusers`` (vendored diffusers 0.3.0) cannot be imported as a package
    because its ``__init__`` chain needs the absent ``diffusers`` distribution; we register empty
    parent packages and load the leaf files (``models/unet_2d_condition.py`` ...) by path.
  * ``utils/utils.py`` imports matplotlib at top level (absent here): stubbed in ``sys.modules``.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")
_MD = os.path.join(REF, "models", "edict", "my_diffusers")
_PKG = "my_diffusers_shim"


def available() -> bool:
    return os.path.isdir(_MD)


def _mkpkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


_loaded = {}


def load_my_diffusers():
    """Returns a namespace with UNet2DConditionModel, DDIMScheduler, CrossAttention of the vendored copy."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    root = _mkpkg(_PKG, _MD)
    root.__version__ = "0.3.0"
    _mkpkg(_PKG + ".utils", os.path.join(_MD, "utils"))
    for leaf in ("import_utils", "logging", "outputs"):
        _load(f"{_PKG}.utils.{leaf}", os.path.join(_MD, "utils", leaf + ".py"))
    utils = _load(_PKG + ".utils", os.path.join(_MD, "utils", "__init__.py"))
    utils.__path__ = [os.path.join(_MD, "utils")]
    _load(_PKG + ".configuration_utils", os.path.join(_MD, "configuration_utils.py"))
    _load(_PKG + ".modeling_utils", os.path.join(_MD, "modeling_utils.py"))
    _mkpkg(_PKG + ".models", os.path.join(_MD, "models"))
    mods = {}
    for leaf in ("embeddings", "attention", "resnet", "unet_blocks", "unet_2d_condition"):
        mods[leaf] = _load(f"{_PKG}.models.{leaf}", os.path.join(_MD, "models", leaf + ".py"))
    _mkpkg(_PKG + ".schedulers", os.path.join(_MD, "schedulers"))
    _load(_PKG + ".schedulers.scheduling_utils", os.path.join(_MD, "schedulers", "scheduling_utils.py"))
    sched = _load(_PKG + ".schedulers.scheduling_ddim", os.path.join(_MD, "schedulers", "scheduling_ddim.py"))
    vae = _load(f"{_PKG}.models.vae", os.path.join(_MD, "models", "vae.py"))
    _loaded.update(
        AutoencoderKL=vae.AutoencoderKL,
        UNet2DConditionModel=mods["unet_2d_condition"].UNet2DConditionModel,
        CrossAttention=mods["attention"].CrossAttention,
        DDIMScheduler=sched.DDIMScheduler,
        attention=mods["attention"],
    )
    return types.SimpleNamespace(**_loaded)


def load_reference_p2p():
    """Imports the reference's own models/p2p/*.py and utils/utils.py unmodified."""
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.__path__ = []
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        mpl.backends = types.ModuleType("matplotlib.backends")
        mpl.backends.__path__ = []
        agg = types.ModuleType("matplotlib.backends.backend_agg")
        agg.FigureCanvasAgg = object
        mpl.backends.backend_agg = agg
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot
        sys.modules["matplotlib.backends"] = mpl.backends
        sys.modules["matplotlib.backends.backend_agg"] = agg
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    ns = types.SimpleNamespace()
    ns.inversion = importlib.import_module("models.p2p.inversion")
    ns.attention_control = importlib.import_module("models.p2p.attention_control")
    ns.p2p_guidance_forward = importlib.import_module("models.p2p.p2p_guidance_forward")
    ns.seq_aligner = importlib.import_module("models.p2p.seq_aligner")
    ns.utils = importlib.import_module("utils.utils")
    return ns


def load_reference_edict(unet, clip, clip_tokenizer, device="cpu"):
    """The reference's EDICT functions (models/edict/edict_functions.py) as a dict of callables, UNMODIFIED.

    The module cannot be imported: at import time it downloads CLIP / the SD UNet / the VAE from the hub and moves them to
    'cuda' (:36-53), and its functions find `unet`, `clip`, `clip_tokenizer`, `device` as module globals.  So the file is
    parsed, only its top-level `def`s are compiled - from the reference's own source text, nothing is copied into this
    repository - and they are executed in a namespace whose globals are the objects passed in (the vendored fp64 UNet
    with synthetic weights, a CLIP stand-in, the whitespace tokenizer).  `coupled_stablediffusion`, `forward_step`,
    `reverse_step`, `init_attention_*`, `use_last_*` / `save_last_*` then run exactly as written.
    """
    import ast
    import math
    import random
    from difflib import SequenceMatcher

    import numpy as np
    import torch
    import torch.nn.functional as F
    from PIL import Image
    from tqdm.auto import tqdm

    md = load_my_diffusers()
    path = os.path.join(REF, "models", "edict", "edict_functions.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    ns = dict(torch=torch, np=np, F=F, math=math, random=random, Image=Image, tqdm=tqdm, SequenceMatcher=SequenceMatcher,
              DDIMScheduler=md.DDIMScheduler, unet=unet, clip=clip, clip_tokenizer=clip_tokenizer, device=device, vae=None)
    exec(compile(tree, path, "exec"), ns)
    return ns
