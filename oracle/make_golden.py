"""TEST INFRASTRUCTURE ONLY -- generates the committed fixtures under tests/golden/ by running the UNMODIFIED reference
(`/root/reference`, through oracle/ref_shim.py) on the seeded synthetic inputs of pnpinversion_b200/synth.py.

Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_golden tables        # integer/0-1 host tables  (seconds)
    python -m oracle.make_golden unet          # single UNet forwards, fp64 vendored UNet (minutes)
    python -m oracle.make_golden pipeline 3    # directinversion+p2p, 3 DDIM steps, full-size UNet (~10 min)
    python -m oracle.make_golden masactrl      # one B=4 forward through the reference's MutualSelfAttentionControl
    python -m oracle.make_golden edict         # EDICT: 2 coupled steps of inversion + 2 of P2P generation (~5 min)
    python -m oracle.make_golden vae           # vendored AutoencoderKL: encode a 64x64 image, decode an 8x8 latent
    python -m oracle.make_golden clip          # transformers CLIPTextModel (the reference's text encoder) on four prompts

What runs is the reference's own `DirectInversion.invert`, `direct_inversion_p2p_guidance_forward`,
`AttentionStore / AttentionRefine / AttentionReweight / LocalBlend`, `register_attention_control`,
`seq_aligner`, `utils.get_word_inds / get_time_words_attention_alpha` against the reference's vendored
`UNet2DConditionModel` + `DDIMScheduler` (diffusers 0.3.0, fp64) loaded with the synthetic fp16-rounded weights.
"""
from __future__ import annotations

import json
import os
import sys
import time
import types

import numpy as np
import torch

from oracle import ref_shim
from pnpinversion_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

PROMPT_PAIRS = [
    list(synth.CAT_PROMPTS),
    ["a cat sitting on a table with a green eyes", "a dog sitting on a table with a green eyes"],
    ["a photo of a house on a hill", "a photo of a red house on a snowy hill at night"],
    ["a round cake with orange frosting on a wooden plate", "a square cake with orange frosting on a wooden plate"],
    ["two birds on a branch", "two colorful birds on a branch"],
    ["a man riding a horse", "a man riding a horse"],
    ["the quick brown fox jumps over the lazy dog", "the quick red fox leaps over the sleepy dog quickly"],
]
BLEND = [("cat", "cat"), ("cat", "dog"), ("house", "house"), ("cake", "cake"), ("birds", "birds"), ("horse", "horse"),
         ("fox", "fox")]
EQ_WORD = ["watercolor", "dog", "red", "square", "colorful", "horse", "red"]


def build_model(dtype=torch.float64):
    md = ref_shim.load_my_diffusers()
    unet = md.UNet2DConditionModel(sample_size=64, cross_attention_dim=768)
    sd = synth.synth_unet_state_dict(0)
    unet.load_state_dict({k: v.to(dtype) for k, v in sd.items()})
    unet = unet.to(dtype).eval()
    sched = md.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                             set_alpha_to_one=False)

    class _Vae:  # image2latent passes 4-D tensors through (utils/utils.py:73-74); decode only feeds the unused image_rec
        def decode(self, z):
            return {"sample": torch.zeros(z.shape[0], 3, 8, 8, dtype=z.dtype)}

    model = types.SimpleNamespace(unet=unet, scheduler=sched, vae=_Vae(), tokenizer=synth.FakeTokenizer(),
                                  text_encoder=synth.SynthTextEncoder(dtype=dtype), device=torch.device("cpu"))
    return model


def make_controller_cpu(ref, model, prompts, num_steps, cross=0.4, self_=0.6, blend_word=None, eq_params=None):
    """make_controller (attention_control.py:366-405) with device='cpu' (it hard-codes 'cuda')."""
    ac = ref.attention_control
    tok = model.tokenizer
    lb = None
    if blend_word is not None:
        lb = ac.LocalBlend(prompts, blend_word, tokenizer=tok, device="cpu", num_ddim_steps=num_steps)
    ctrl = ac.AttentionRefine(prompts, num_steps, cross_replace_steps={"default_": cross}, self_replace_steps=self_,
                              local_blend=lb, tokenizer=tok, device="cpu")
    if eq_params is not None:
        eq = ac.get_equalizer(prompts[1], eq_params["words"], eq_params["values"], tokenizer=tok)
        ctrl = ac.AttentionReweight(prompts, num_steps, cross_replace_steps={"default_": cross},
                                    self_replace_steps=self_, equalizer=eq, local_blend=lb, controller=ctrl,
                                    device="cpu")
    return ctrl


def gen_tables():
    ref = ref_shim.load_reference_p2p()
    tok = synth.FakeTokenizer()
    out = []
    for (src, tgt), (bs, bt), eqw in zip(PROMPT_PAIRS, BLEND, EQ_WORD):
        prompts = [src, tgt]
        mapper, alphas = ref.seq_aligner.get_refinement_mapper(prompts, tok)
        entry = {"prompts": prompts, "mapper": mapper[0].tolist(), "alphas": alphas[0].tolist()}
        for n in (50, 20, 3):
            a = ref.utils.get_time_words_attention_alpha(prompts, n, {"default_": 0.4}, tok)
            entry[f"cross_alpha_{n}"] = a.reshape(n + 1, 77)[:, 0].tolist()
        entry["inds_src"] = ref.utils.get_word_inds(src, bs, tok).tolist()
        entry["inds_tgt"] = ref.utils.get_word_inds(tgt, bt, tok).tolist()
        entry["blend"] = [bs, bt]
        entry["eq_word"] = eqw
        entry["equalizer"] = ref.attention_control.get_equalizer(tgt, (eqw,), (2,), tok)[0].tolist()
        if len(src.split(" ")) == len(tgt.split(" ")):
            entry["replace_mapper"] = ref.seq_aligner.get_replacement_mapper(prompts, tok)[0].tolist()
        out.append(entry)
    # schedule tables of the vendored scheduler (float64) for 50 / 20 / 3 steps
    md = ref_shim.load_my_diffusers()
    sched = md.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                             set_alpha_to_one=False)
    sch = {"alphas_cumprod_f64_first_last": [float(sched.alphas_cumprod[0]), float(sched.alphas_cumprod[-1])],
           "alphas_cumprod_f64_every_100": [float(sched.alphas_cumprod[i]) for i in range(0, 1000, 100)],
           "final_alpha_cumprod": float(sched.final_alpha_cumprod)}
    for n in (50, 20, 3):
        sched.set_timesteps(n)
        sch[f"timesteps_{n}"] = [int(t) for t in sched.timesteps]
    with open(os.path.join(GOLD, "tables.json"), "w") as f:
        json.dump({"pairs": out, "schedule": sch}, f)
    print("tables.json written:", len(out), "pairs")


def _ctx(model, prompts):
    tok, te = model.tokenizer, model.text_encoder
    un = te(tok([""] * len(prompts)).input_ids)[0]
    tx = te(tok(prompts).input_ids)[0]
    return torch.cat([un, tx])


def gen_unet():
    ref = ref_shim.load_reference_p2p()
    model = build_model()
    prompts = list(synth.CAT_PROMPTS)
    ctx = _ctx(model, prompts)  # [uncond, uncond, src, tgt]
    res = {}
    with torch.no_grad():
        # (a) B=1, no controller, the ddim_loop call shape (inversion.py:272-274,315-316)
        ref.attention_control.register_attention_control(model, None)
        x = synth.synth_latent(0).double()
        t0 = time.time()
        res["a_eps"] = model.unet(x, torch.tensor(981), encoder_hidden_states=ctx[2:3])["sample"].float().numpy()
        print("case a", time.time() - t0)
        # (b) B=4, Refine+Reweight+LocalBlend controller at step 0 (cross gate 1, self-replace on)
        lat = torch.cat([synth.synth_latent(0), synth.synth_latent(1)]).double()
        for name, step in (("b", 0), ("c", 35)):
            ctrl = make_controller_cpu(ref, model, prompts, 50, blend_word=(("cat",), ("cat",)),
                                       eq_params={"words": ("watercolor",), "values": (2,)})
            ctrl.cur_step = step
            ref.attention_control.register_attention_control(model, ctrl)
            t0 = time.time()
            eps = model.unet(torch.cat([lat] * 2), torch.tensor(601), encoder_hidden_states=ctx)["sample"]
            print("case", name, time.time() - t0, "cur_step after", ctrl.cur_step)
            res[f"{name}_eps"] = eps.float().numpy()
            # what LocalBlend would read after this single step: the five 16x16 maps, reduced like get_mask does
            maps = ctrl.attention_store["down_cross"][2:4] + ctrl.attention_store["up_cross"][:3]
            maps = torch.cat([m.reshape(2, -1, 1, 16, 16, 77) for m in maps], dim=1)
            res[f"{name}_maps_mean"] = maps.mean(1).reshape(2, 16, 16, 77).float().numpy()  # (2,16,16,77)
    np.savez_compressed(os.path.join(GOLD, "unet_forward.npz"), **res)
    print("unet_forward.npz written")


def gen_pipeline(n_steps: int):
    ref = ref_shim.load_reference_p2p()
    model = build_model()
    prompts = list(synth.CAT_PROMPTS)
    z0 = synth.synth_latent(0).double()
    t0 = time.time()
    inv = ref.inversion.DirectInversion(model=model, num_ddim_steps=n_steps)
    model.scheduler.set_timesteps(n_steps)
    _, _, x_stars, noise_loss = inv.invert(image_gt=z0, prompt=prompts, guidance_scale=7.5)
    print("invert done", time.time() - t0)
    x_t = x_stars[-1]
    ctrl = ref.attention_control.AttentionStore()
    recon, _ = ref.p2p_guidance_forward.direct_inversion_p2p_guidance_forward(
        model=model, prompt=prompts, controller=ctrl, noise_loss_list=noise_loss, latent=x_t,
        num_inference_steps=n_steps, guidance_scale=7.5, generator=None)
    print("recon done", time.time() - t0)
    ctrl = make_controller_cpu(ref, model, prompts, n_steps, blend_word=(("cat",), ("cat",)),
                               eq_params={"words": ("watercolor",), "values": (2,)})
    edit, _ = ref.p2p_guidance_forward.direct_inversion_p2p_guidance_forward(
        model=model, prompt=prompts, controller=ctrl, noise_loss_list=noise_loss, latent=x_t,
        num_inference_steps=n_steps, guidance_scale=7.5, generator=None)
    print("edit done", time.time() - t0)
    np.savez_compressed(
        os.path.join(GOLD, f"pipeline_{n_steps}steps.npz"),
        x_stars=torch.cat(x_stars).float().numpy(), noise_loss=torch.stack(noise_loss).float().numpy(),
        recon=recon.float().numpy(), edit=edit.float().numpy())
    print("pipeline fixture written")


def gen_pipeline_full(n_steps: int = 50):
    """BASELINE config 2 at its true length: the reference's own DirectInversion.invert +
    direct_inversion_p2p_guidance_forward (AttentionStore pass, then Refine+Reweight+LocalBlend) over `n_steps` DDIM
    steps on the vendored fp64 UNet = 13 * n_steps sample-forwards (hours of CPU at 50).  Checkpoints after each phase
    under gpurun_out/ (scratch) so a lost session does not lose the run; the committed fixture keeps every x_star, every
    fifth noise_loss, and the two final latent pairs.  Also config 1: a 20-step ddim_loop of the same latent."""
    ref = ref_shim.load_reference_p2p()
    model = build_model()
    prompts = list(synth.CAT_PROMPTS)
    z0 = synth.synth_latent(0).double()
    scratch = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
    os.makedirs(scratch, exist_ok=True)
    t0 = time.time()
    # config 1: 20-step inversion only (inversion.py:308-319 through invert's first half)
    inv20 = ref.inversion.DirectInversion(model=model, num_ddim_steps=20)
    model.scheduler.set_timesteps(20)
    inv20.init_prompt(prompts)  # ddim_loop uses cond_embeddings[[0]] = the source prompt (inversion.py:309-311)
    ref.attention_control.register_attention_control(model, None)
    x20 = inv20.ddim_loop(z0)
    np.savez_compressed(os.path.join(GOLD, "inversion_20steps.npz"), x_stars=torch.cat(x20).float().numpy())
    print("config-1 20-step inversion done", time.time() - t0, flush=True)
    inv = ref.inversion.DirectInversion(model=model, num_ddim_steps=n_steps)
    model.scheduler.set_timesteps(n_steps)
    _, _, x_stars, noise_loss = inv.invert(image_gt=z0, prompt=prompts, guidance_scale=7.5)
    print("invert done", time.time() - t0, flush=True)
    np.savez_compressed(os.path.join(scratch, f"pipeline_{n_steps}_partial_invert.npz"),
                        x_stars=torch.cat(x_stars).numpy(), noise_loss=torch.stack(noise_loss).numpy())
    x_t = x_stars[-1]
    ctrl = ref.attention_control.AttentionStore()
    recon, _ = ref.p2p_guidance_forward.direct_inversion_p2p_guidance_forward(
        model=model, prompt=prompts, controller=ctrl, noise_loss_list=noise_loss, latent=x_t,
        num_inference_steps=n_steps, guidance_scale=7.5, generator=None)
    print("recon done", time.time() - t0, flush=True)
    np.savez_compressed(os.path.join(scratch, f"pipeline_{n_steps}_partial_recon.npz"), recon=recon.numpy())
    ctrl = make_controller_cpu(ref, model, prompts, n_steps, blend_word=(("cat",), ("cat",)),
                               eq_params={"words": ("watercolor",), "values": (2,)})
    edit, _ = ref.p2p_guidance_forward.direct_inversion_p2p_guidance_forward(
        model=model, prompt=prompts, controller=ctrl, noise_loss_list=noise_loss, latent=x_t,
        num_inference_steps=n_steps, guidance_scale=7.5, generator=None)
    print("edit done", time.time() - t0, flush=True)
    keep = sorted(set(list(range(0, n_steps, 5)) + [n_steps - 1]))
    np.savez_compressed(
        os.path.join(GOLD, f"pipeline_{n_steps}steps.npz"),
        x_stars=torch.cat(x_stars).float().numpy(), noise_loss_idx=np.array(keep, np.int64),
        noise_loss=torch.stack([noise_loss[i] for i in keep]).float().numpy(),
        recon=recon.float().numpy(), edit=edit.float().numpy())
    print("pipeline fixture written", flush=True)


def gen_masactrl():
    """One B=4 UNet forward with the reference's own MutualSelfAttentionControl (models/masactrl/masactrl.py:14-72)
    registered through its own regiter_attention_editor_diffusers (masactrl_utils.py:79-144) on the vendored UNet.
    The registration matches the class name 'Attention' (diffusers >= 0.15); the vendored class is named
    CrossAttention, so the harness renames it -- no reference code is modified."""
    import importlib
    import sys as _sys

    if ref_shim.REF not in _sys.path:
        _sys.path.insert(0, ref_shim.REF)
    masactrl = importlib.import_module("models.masactrl.masactrl")
    mutils = importlib.import_module("models.masactrl.masactrl_utils")
    model = build_model()
    md = ref_shim.load_my_diffusers()
    md.CrossAttention.__name__ = "Attention"
    prompts = ["", synth.CAT_PROMPTS[1]]
    ctx = _ctx(model, prompts)
    lat = torch.cat([synth.synth_latent(0), synth.synth_latent(1)]).double()
    res = {}
    with torch.no_grad():
        for name, step in (("on", 10), ("off", 2)):
            editor = masactrl.MutualSelfAttentionControl(4, 10)
            mutils.regiter_attention_editor_diffusers(model, editor)
            assert editor.num_att_layers == 32
            editor.cur_step = step
            t0 = time.time()
            eps = model.unet(torch.cat([lat] * 2), torch.tensor(401), encoder_hidden_states=ctx)["sample"]
            print("masactrl", name, time.time() - t0, "cur_step after", editor.cur_step)
            res[f"{name}_eps"] = eps.float().numpy()
    md.CrossAttention.__name__ = "CrossAttention"
    np.savez_compressed(os.path.join(GOLD, "masactrl_forward.npz"), **res)
    print("masactrl_forward.npz written")


def gen_masactrl_pipeline(n_steps: int = 4):
    """`directinversion+masactrl` end to end with the reference's own loops (run_editing_masactrl.py:89-129):
    DirectInversion.invert with prompts ["", target] (models/p2p/inversion.py), then MasaCtrlPipeline.__call__
    (models/masactrl/diffuser_utils.py:90-193) twice - direct synthesis with the target prompt, and the mutual
    self-attention pass with the rectification of :183-184 - on the vendored fp64 UNet.  MasaCtrlPipeline subclasses the
    absent diffusers StableDiffusionPipeline, so its methods are compiled from the reference's source text (nothing is
    copied into this repository) and bound to a harness object carrying unet / scheduler / tokenizer / text_encoder; the
    harness's latent2image returns the latents themselves (the VAE is a separate fixture)."""
    import ast
    import importlib
    import sys as _sys

    from tqdm import tqdm

    ref = ref_shim.load_reference_p2p()
    if ref_shim.REF not in _sys.path:
        _sys.path.insert(0, ref_shim.REF)
    masactrl = importlib.import_module("models.masactrl.masactrl")
    mutils = importlib.import_module("models.masactrl.masactrl_utils")
    model = build_model()
    md = ref_shim.load_my_diffusers()
    path = os.path.join(ref_shim.REF, "models", "masactrl", "diffuser_utils.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MasaCtrlPipeline")
    funcs = ast.Module(body=[n for n in cls.body if isinstance(n, ast.FunctionDef)], type_ignores=[])
    ns = dict(torch=torch, np=np, tqdm=tqdm)
    exec(compile(funcs, path, "exec"), ns)

    class Harness:
        next_step, step, invert = ns["next_step"], ns["step"], ns["invert"]
        __call__ = ns["__call__"]

        def latent2image(self, latents, return_type="pt"):
            return latents

    assert 1000 % n_steps == 0, "the vendored 0.3.0 scheduler yields n + 1 timesteps unless n divides 1000"
    pipe = Harness()
    pipe.unet, pipe.scheduler, pipe.tokenizer, pipe.text_encoder = model.unet, model.scheduler, model.tokenizer, model.text_encoder
    tgt = synth.CAT_PROMPTS[1]
    prompts = ["", tgt]
    z0 = synth.synth_latent(3).double()
    t0 = time.time()
    model.scheduler.set_timesteps(n_steps)
    inv = ref.inversion.DirectInversion(model=model, num_ddim_steps=n_steps)
    _, _, x_stars, noise_loss = inv.invert(image_gt=z0, prompt=prompts, guidance_scale=7.5)
    print("invert done", time.time() - t0, flush=True)
    x_t = x_stars[-1]
    md.CrossAttention.__name__ = "Attention"  # masactrl_utils.py:129 matches the diffusers >= 0.15 class name
    try:
        mutils.regiter_attention_editor_diffusers(model, mutils.AttentionBase())
        fixed = pipe([tgt], latents=x_t, num_inference_steps=n_steps, guidance_scale=7.5, noise_loss_list=None)
        print("direct synthesis done", time.time() - t0, flush=True)
        # run_editing_masactrl.py:89 defaults (step 4, layer 10) for a full schedule; step 1 for the short fixtures
        editor = masactrl.MutualSelfAttentionControl(4 if n_steps >= 10 else 1, 10, total_steps=n_steps)
        mutils.regiter_attention_editor_diffusers(model, editor)
        out = pipe(prompts, latents=x_t.expand(2, -1, -1, -1), num_inference_steps=n_steps, guidance_scale=7.5,
                   noise_loss_list=noise_loss)
        print("masactrl pass done", time.time() - t0, flush=True)
    finally:
        md.CrossAttention.__name__ = "CrossAttention"
    xs, nl = torch.cat(x_stars).float().numpy(), torch.stack(noise_loss).float().numpy()
    xi, ni = np.arange(len(xs)), np.arange(len(nl))
    if n_steps > 10:  # full schedules: every fifth entry (and the last) keeps the fixture small
        xi = np.array(sorted(set(list(range(0, len(xs), 5)) + [len(xs) - 1])))
        ni = np.array(sorted(set(list(range(0, len(nl), 5)) + [len(nl) - 1])))
    np.savez_compressed(os.path.join(GOLD, f"masactrl_pipeline_{n_steps}steps.npz"), x_stars=xs[xi], x_index=xi,
                        noise_loss=nl[ni], noise_index=ni, fixed=fixed.float().numpy(), out=out.float().numpy())
    print("masactrl pipeline fixture written")


def gen_pnp(n_steps: int = 4):
    """Plug-and-Play features with the PnP-Inversion source branch, by the reference's own functions
    (run_editing_pnp.py): `Preprocess.ddim_inversion` / `ddim_sample` (:88-134), `register_time` (:150-174),
    `register_attention_control_efficient` (:176-242), `register_conv_control_efficient` (:244-294) and
    `PNP.denoise_step` (:344-361), compiled from the reference's source text (the module itself loads diffusers models at
    import) and run on the vendored fp64 UNet.  Harness adaptations, none of them in reference code: the vendored
    CrossAttention's head reshapes get the diffusers >= 0.10 names the patched forward calls, the scheduler's timesteps
    are a tensor with steps_offset 1 (the runwayml/stable-diffusion-v1-5 scheduler config)."""
    import ast

    model = build_model()
    md = ref_shim.load_my_diffusers()
    path = os.path.join(ref_shim.REF, "run_editing_pnp.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    want_fn = {"register_time", "register_attention_control_efficient", "register_conv_control_efficient"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want_fn]
    for cls_name, methods in (("Preprocess", {"ddim_inversion", "ddim_sample"}), ("PNP", {"denoise_step"})):
        cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
        body += [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in methods]
    ns = dict(torch=torch, np=np)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    for m in model.unet.modules():
        if type(m).__name__ == "CrossAttention":
            m.head_to_batch_dim = m.reshape_heads_to_batch_dim
            m.batch_to_head_dim = m.reshape_batch_dim_to_heads
    assert 1000 % n_steps == 0, "the vendored 0.3.0 scheduler yields n + 1 timesteps unless n divides 1000"
    sched = model.scheduler
    sched.set_timesteps(n_steps, offset=1)
    sched.timesteps = torch.from_numpy(np.ascontiguousarray(sched.timesteps)).long()
    src, tgt = synth.CAT_PROMPTS
    tok, te = model.tokenizer, model.text_encoder
    emb = lambda p: te(tok([p]).input_ids)[0]
    pre = types.SimpleNamespace(unet=model.unet, scheduler=sched)
    z0 = synth.synth_latent(8).double()
    t0 = time.time()
    with torch.no_grad():
        inverted_x = ns["ddim_inversion"](pre, emb(src), z0)
        rec = ns["ddim_sample"](pre, inverted_x[-1], emb(src))
        print("inversion + reconstruction done", time.time() - t0, flush=True)
        pnp = types.SimpleNamespace(unet=model.unet, scheduler=sched, pnp_guidance_embeds=emb(""),
                                    text_embeds=torch.cat([emb("ugly, blurry, black, low res, unrealistic"), emb(tgt)]))
        qk_t = sched.timesteps[: int(n_steps * 0.5)]
        conv_t = sched.timesteps[: int(n_steps * 0.8)]
        ns["register_attention_control_efficient"](pnp, qk_t)
        ns["register_conv_control_efficient"](pnp, conv_t)
        x = inverted_x[-1]
        xs = []
        for i, t in enumerate(sched.timesteps):
            x = ns["denoise_step"](pnp, x, t, 7.5, inverted_x[-1 - i])
            xs.append(x)
        print("pnp sampling done", time.time() - t0, flush=True)
    np.savez_compressed(os.path.join(GOLD, f"pnp_features_{n_steps}steps.npz"),
                        timesteps=sched.timesteps.numpy(), inverted_x=torch.cat(inverted_x).float().numpy(),
                        rec=torch.cat(rec).float().numpy(), xs=torch.cat(xs).float().numpy(),
                        qk_t=qk_t.numpy(), conv_t=conv_t.numpy())
    print("pnp fixture written", [int(t) for t in sched.timesteps], "qk", qk_t.tolist(), "conv", conv_t.tolist())


def gen_edict(strength: float = 0.04):
    """strength 0.04: the two-step fixture described below; any other value (0.8 = the reference's default,
    run_editing_edict.py:43) runs the FULL schedule - 40 noising + 40 generation steps of 50 - and stores only the start
    latent, the noised pair and the edited pair (tests/golden/edict_<n>steps.npz, ~400 fp64 UNet sample-forwards).

    The reference's own `coupled_stablediffusion` (models/edict/edict_functions.py:707-956): deterministic noising of a
    latent pair over the last two of 50 timesteps (init_image_strength 0.04 -> t = 0, 20), then generation from that pair
    with the Prompt-to-Prompt attention reuse (prompt_edit given).  Every UNet call is recorded (timestep, which text
    embedding, checksum of the input latent, the predicted noise) so that the CPU restatement can be replayed against the
    exact call sequence without running a UNet, and pinned on single calls where it does."""
    model = build_model()
    tok, te = model.tokenizer, model.text_encoder
    src, tgt = synth.CAT_PROMPTS

    class Tok:  # CLIPTokenizer call surface used by edict_functions.py:818-838
        model_max_length = tok.model_max_length

        def __call__(self, text, padding="max_length", max_length=77, truncation=True, return_tensors="pt",
                     return_overflowing_tokens=True):
            return tok(text, padding=padding, max_length=max_length, truncation=truncation, return_tensors=return_tensors)

    embs, keep = {}, []

    class Clip:  # clip(ids).last_hidden_state
        def __call__(self, ids):
            e = te(ids)[0]
            keep.append(e)  # keeps id(e) unique for the lifetime of the run
            embs[id(e)] = self.n  # 0 = null prompt, 1 = prompt, 2 = prompt_edit (order of edict_functions.py:818-838)
            self.n += 1
            return types.SimpleNamespace(last_hidden_state=e)

    clip = Clip()
    clip.n = 0

    calls = []

    class Recorder:  # the `unet` global of the reference functions
        in_channels = model.unet.in_channels

        def named_modules(self):
            return model.unet.named_modules()

        def __call__(self, x, t, encoder_hidden_states=None):
            out = model.unet(x, t, encoder_hidden_states=encoder_hidden_states)
            calls.append(dict(t=int(t), ctx=embs[id(encoder_hidden_states)], s=float(x.sum()), a=float(x.abs().sum()),
                              eps=out.sample.detach().clone()))
            return out

    ns = ref_shim.load_reference_edict(Recorder(), clip, Tok(), "cpu")
    z = synth.synth_latent(6).double()
    t0 = time.time()
    full = abs(strength - 0.04) > 1e-9
    lat = ns["coupled_stablediffusion"](src, reverse=True, init_image=[z, z.clone()], init_image_strength=strength, steps=50,
                                        mix_weight=0.93, guidance_scale=3.0)
    n_rev = len(calls)
    print(f"reverse pass: {n_rev} UNet calls, {time.time() - t0:.0f}s", flush=True)
    clip.n = 0  # the generation pass encodes null prompt, prompt, prompt_edit again
    out = ns["coupled_stablediffusion"](src, tgt, fixed_starting_latent=lat, init_image_strength=strength, steps=50,
                                        mix_weight=0.93, guidance_scale=3.0, return_latents=True)
    print(f"forward P2P pass: {len(calls) - n_rev} UNet calls, {time.time() - t0:.0f}s", flush=True)
    if full:
        n = int(50 * strength)
        np.savez_compressed(os.path.join(GOLD, f"edict_{n}steps.npz"), z=z.numpy().astype(np.float32),
                            lat=torch.stack(lat).float().numpy(), out=torch.stack(out).float().numpy(),
                            strength=np.float64(strength), n_reverse_calls=np.int64(n_rev), n_calls=np.int64(len(calls)))
        print(f"wrote edict_{n}steps.npz")
        return
    attn2 = next(m for n, m in model.unet.named_modules() if type(m).__name__ == "CrossAttention" and "attn2" in n)
    np.savez_compressed(
        os.path.join(GOLD, "edict_2steps.npz"), z=z.numpy().astype(np.float32),
        lat=torch.stack(lat).numpy(), out=torch.stack(out).numpy(), n_reverse_calls=np.int64(n_rev),
        call_t=np.array([c["t"] for c in calls], np.int64), call_ctx=np.array([c["ctx"] for c in calls], np.int64),
        call_in_sum=np.array([c["s"] for c in calls], np.float64), call_in_abs=np.array([c["a"] for c in calls], np.float64),
        call_eps=torch.cat([c["eps"] for c in calls]).numpy().astype(np.float32),
        edit_mask=attn2.last_attn_slice_mask.numpy(), edit_indices=attn2.last_attn_slice_indices.numpy())
    print("wrote edict_2steps.npz:", [(c["t"], c["ctx"]) for c in calls])


def gen_vae():
    """The vendored AutoencoderKL (models/edict/my_diffusers/models/vae.py:480-557, SD-1.x configuration, fp64) on the
    synthetic VAE weights: posterior moments of a seeded 64x64 image and the decode of a seeded 8x8 latent - the two calls
    of utils/utils.py:58-80 at a size that keeps the fixture small (the network is fully convolutional apart from the
    single-head attention of the mid blocks, whose token count changes with the size but not its arithmetic)."""
    md = ref_shim.load_my_diffusers()
    vae = md.AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                           up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512),
                           layers_per_block=2, act_fn="silu", latent_channels=4, sample_size=512)
    sd = synth.synth_vae_state_dict(0)
    vae.load_state_dict({k: v.double() for k, v in sd.items()})
    vae = vae.double().eval()
    g = torch.Generator().manual_seed(4242)
    img = (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).double()
    z = torch.randn(1, 4, 8, 8, generator=g).double()
    dist = vae.encode(img).latent_dist
    dec = vae.decode(z).sample
    np.savez_compressed(os.path.join(GOLD, "vae_small.npz"), img=img.numpy().astype(np.float32),
                        z=z.numpy().astype(np.float32), mean=dist.mean.numpy(), logvar=dist.logvar.numpy(),
                        dec=dec.numpy())
    print("wrote vae_small.npz: mean", tuple(dist.mean.shape), "dec", tuple(dec.shape),
          "mean rms %.3f dec rms %.3f" % (float(dist.mean.pow(2).mean().sqrt()), float(dec.pow(2).mean().sqrt())))


def gen_clip():
    """The text encoder the reference calls (`model.text_encoder(ids)[0]`, models/p2p/inversion.py:42,50): the installed
    `transformers.CLIPTextModel` itself (SD-1.x text tower configuration, fp64) on the synthetic weights of
    synth.synth_clip_state_dict and the token ids of the null prompt, the cat prompt pair and a 77-token prompt that is
    truncated (no padding at all)."""
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel

    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    m = CLIPTextModel(cfg).double().eval()
    sd = synth.synth_clip_state_dict(0)
    missing, unexpected = m.load_state_dict({k: v.double() for k, v in sd.items()}, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    tok = synth.FakeTokenizer()
    ids = tok(["", synth.CAT_PROMPTS[0], synth.CAT_PROMPTS[1], " ".join(f"w{i}" for i in range(90))]).input_ids
    out = m(ids)[0]
    np.savez_compressed(os.path.join(GOLD, "clip_text.npz"), ids=ids.numpy().astype(np.int32),
                        out=out.numpy().astype(np.float32), transformers_version=np.array(transformers.__version__))
    print("wrote clip_text.npz:", tuple(out.shape), "rms %.3f" % float(out.pow(2).mean().sqrt()), "transformers", transformers.__version__)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(False)
    what = sys.argv[1]
    if what == "tables":
        gen_tables()
    elif what == "unet":
        gen_unet()
    elif what == "pipeline":
        gen_pipeline(int(sys.argv[2]))
    elif what == "pipeline_full":
        gen_pipeline_full(int(sys.argv[2]) if len(sys.argv) > 2 else 50)
    elif what == "masactrl_pipeline":
        gen_masactrl_pipeline(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
    elif what == "pnp":
        gen_pnp(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
    elif what == "masactrl":
        gen_masactrl()
    elif what == "edict":
        gen_edict(float(sys.argv[2]) if len(sys.argv) > 2 else 0.04)
    elif what == "vae":
        gen_vae()
    elif what == "clip":
        gen_clip()
