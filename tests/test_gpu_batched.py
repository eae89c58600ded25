"""The C-side step loops (`pnp_run_loop`), image batching, shared-weight clones, the span mapper of fractional
AttentionReplace and the batched LocalBlend (with substruct words) -- all through the C ABI on a B200."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from pnpinversion_b200 import _lib, synth
from pnpinversion_b200.model import FusedModel
from pnpinversion_b200.p2p_editor import P2PEditor
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
H = 8
BLEND = (("cat",), ("cat",))
EQ = {"words": ("watercolor",), "values": (2,)}


@pytest.fixture(scope="module")
def model(cuda):
    m = FusedModel.synthetic(device="cuda:0", max_batch=16)
    yield m
    m.unet.close()


def _edit(model, method, fused, steps=4, z=None, **kw):
    ed = P2PEditor([method], "cuda:0", num_ddim_steps=steps, model=model, fused_loops=fused)
    src, tgt = synth.CAT_PROMPTS
    r = ed(method, image_path=synth.synth_latent(0) if z is None else z, prompt_src=src, prompt_tar=tgt,
           cross_replace_steps=0.4, self_replace_steps=0.6, blend_word=BLEND, eq_params=EQ, **kw)
    torch.cuda.synchronize()
    return r


@pytest.mark.parametrize("method", ["directinversion+p2p", "ablation_directinversion_add-target+p2p",
                                    "ablation_directinversion_add-source+p2p", "ablation_directinversion_08+p2p",
                                    "ablation_directinversion_interval_2+p2p", "directinversion+p2p_guidance_25_75",
                                    "directinversion+p2p_guidance_0_5"])
def test_c_loops_are_bit_identical_to_the_python_loops(model, method):
    """`pnp_run_loop` (INVERT / OFFSET / FORWARD incl. per-step controller descriptors and LocalBlend) against the Python
    loops that mirror the reference's (inversion.py, p2p_guidance_forward.py): same kernels, same order, same bits."""
    a = _edit(model, method, fused=True)
    b = _edit(model, method, fused=False)
    if method.endswith("guidance_0_5"):
        # inverse guidance 0: the Python loop evaluates u + 0 * (c - u) from a B=2 call like the reference, the C loop
        # issues the unconditional row alone (B=1): same arithmetic, possibly another GEMM tiling -> rounding-level
        ea, eb = G.rel_l2(a.x_stars[-1], b.x_stars[-1]), G.rel_l2(a.latents[1], b.latents[1])
        print(f"inverse guidance 0: C loop (B=1 unconditional) vs Python loop (B=2 CFG with g=0): x_T {ea:.2e} edit {eb:.2e}")
        assert ea < 8e-3 and eb < 0.25  # 4 steps of CFG on two rounding-noise realisations
        return
    for xa, xb in zip(a.x_stars, b.x_stars):
        assert torch.equal(xa, xb)
    for la, lb in zip(a.noise_loss_list, b.noise_loss_list):
        assert torch.equal(la, lb)
    assert torch.equal(a.reconstruct_latent, b.reconstruct_latent)
    assert torch.equal(a.latents, b.latents)
    assert torch.isfinite(a.latents).all()
    if "interval_2" in method:  # the skipped steps carry a zero offset (inversion.py:514-517)
        assert float(a.noise_loss_list[1].abs().max()) == 0.0 and float(a.noise_loss_list[0].abs().max()) > 0.0


def test_image_batch_matches_the_single_image_runs(model):
    """Three images through ONE pass (UNet batch 3 / 12) against each image alone (batch 1 / 4).  Tile / split-K choices
    may differ with the batch size, so the comparison is to rounding of the fp16 pipeline, not to the bit."""
    ed = P2PEditor(["directinversion+p2p"], "cuda:0", num_ddim_steps=4, model=model)
    src, tgt = synth.CAT_PROMPTS
    zs = torch.cat([synth.synth_latent(i) for i in range(3)]).cuda()
    pairs_src = [src, "a photo of a house on a hill", src]
    pairs_tgt = [tgt, "a photo of a red house on a snowy hill at night", tgt]
    blends = [BLEND, (("house",), ("house",)), BLEND]
    eqs = [EQ, {"words": ("red",), "values": (2,)}, EQ]
    res = ed.edit_batch(zs, pairs_src, pairs_tgt, blend_word=blends, eq_params=eqs, per_image_params=True)
    torch.cuda.synchronize()
    assert torch.isfinite(res.latents).all()
    for i in range(3):
        one = ed("directinversion+p2p", image_path=zs[i:i + 1], prompt_src=pairs_src[i], prompt_tar=pairs_tgt[i],
                 blend_word=blends[i], eq_params=eqs[i])
        torch.cuda.synchronize()
        x_stars, nl, rec, lat = res.image(i)
        ex = G.rel_l2(x_stars[-1], one.x_stars[-1])
        print(f"image {i}: batched vs single x_T rel-L2 {ex:.3e}")
        assert ex < 8e-3  # 4 inversion steps; both runs sit ~3e-3 from the fp64 reference (tests/test_gpu_pipeline.py)
        # the rectified source branch lands on z0 in the batch as well
        assert (lat[0] - zs[i]).abs().max() < 2e-5 and (rec[0] - zs[i]).abs().max() < 2e-5
        e = G.rel_l2(lat[1], one.latents[1])
        print(f"image {i}: batched vs single edit rel-L2 {e:.3e}")
        # another batch size = another realisation of the rounding noise, amplified by CFG 7.5 over the steps
        assert e < 0.3
    # images 0 and 2 share prompts but not latents: different results; image 0 twice would be identical
    assert G.rel_l2(res.latents[3], res.latents[5]) > 1e-2


def test_unet_rows_are_independent_of_the_batch_size(model):
    """B = 8 / 12 / 16 forwards (MasaCtrl 4-image batches, EDICT B=8) reproduce the B = 4 rows."""
    tok, te = model.tokenizer, model.text_encoder
    prompts = list(synth.CAT_PROMPTS)
    ctx4 = torch.cat([te(tok([""] * 2).input_ids)[0], te(tok(prompts).input_ids)[0]]).cuda().float()
    x4 = torch.cat([synth.synth_latent(i) for i in range(4)]).cuda()
    model.unet.set_controller(None)
    ref = model.unet(x4, 481, encoder_hidden_states=ctx4.contiguous())["sample"]
    for reps in (2, 3, 4):
        xb = torch.cat([x4] * reps).contiguous()
        cb = torch.cat([ctx4] * reps).contiguous()
        out = model.unet(xb, 481, encoder_hidden_states=cb)["sample"]
        torch.cuda.synchronize()
        for r in range(reps):
            e = G.rel_l2(out[4 * r:4 * r + 4], ref)
            print(f"B={4 * reps} rows {4 * r}..{4 * r + 3} vs B=4: rel-L2 {e:.2e}")
            assert e < 6e-3, (reps, r, e)  # each batch size is ~3.3e-3 from the fp64 reference (tools/diag_batch.py)


def test_clone_shares_weights_and_reproduces_the_parent(model):
    clone = model.clone(max_batch=4)
    a = _edit(model, "directinversion+p2p", fused=True)
    b = _edit(clone, "directinversion+p2p", fused=True)
    assert torch.equal(a.latents, b.latents) and torch.equal(a.x_stars[-1], b.x_stars[-1])
    clone.unet.close()
    c = _edit(model, "directinversion+p2p", fused=True)  # the parent is intact after the clone is destroyed
    assert torch.equal(a.latents, c.latents)


def _heads(t, d):
    B, N, _ = t.shape
    return t.reshape(B, N, H, d).permute(0, 2, 1, 3).float()


def test_cross_attention_span_mapper_equals_the_dense_einsum(cuda):
    """AttentionReplace with unequal token spans: the kernel's `weight * sum of count consecutive source tokens` against
    the reference's einsum('hpw,bwn->bhpn', attn_base, mapper) (attention_control.py:303-304) in fp32."""
    from pnpinversion_b200 import seq_aligner
    from pnpinversion_b200.attention_control import AttentionReplace
    tok = synth.PieceTokenizer()
    prompts = ["a watercolor of a cat and a bird", "a pic of a crocodile and a bird"]
    mapper = seq_aligner.get_replacement_mapper(prompts, tok)[0]  # (77,77) with 1/len(target) columns
    start, count, weight = AttentionReplace._columns(mapper)
    assert max(count) > 1 and any(0 < w < 1 for w in weight)
    lib = _lib.load()
    B, N, d = 4, 256, 160
    g = torch.Generator().manual_seed(21)
    q = (torch.randn(B, N, H * d, generator=g) * 1.5).half().to(cuda)
    kv = (torch.randn(B, 77, 2 * H * d, generator=g) * 1.5).half().to(cuda)
    ca = (torch.rand(77, generator=g) > 0.3).float()
    ctrl = _lib.new_ctrl()
    ctrl.cross_base_row[3] = 2
    ctrl.cross_slot[3] = 0
    ctrl.mapper[0][:] = start
    ctrl.map_count[0][:] = count
    ctrl.map_weight[0][:] = weight
    ctrl.cross_alpha[0][:] = ca.tolist()
    out = torch.empty(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_cross_attention(G.ptr(q), G.ptr(kv), B, H, N, d, 77, C.byref(ctrl), None, G.ptr(out),
                                            G.stream()))
    torch.cuda.synchronize()
    c = H * d
    qh, kh, vh = _heads(q, d), _heads(kv[..., :c], d), _heads(kv[..., c:], d)
    p = (qh @ kh.transpose(-1, -2) * d ** -0.5).softmax(-1)
    p = p.clone()
    rep = torch.einsum("hpw,wn->hpn", p[2], mapper.to(cuda))
    cad = ca.to(cuda)
    p[3] = rep * cad + (1 - cad) * p[3]
    ref = (p @ vh).permute(0, 2, 1, 3).reshape(B, N, c)
    assert G.rel_l2(out, ref) < 2e-3
    assert G.rel_l2(out[3], ref[3]) < 2e-3


def test_local_blend_batch_with_substruct_words_vs_oracle(model, cuda):
    """Two (source, target) pairs in one launch, the second with substruct words, against the reference's LocalBlend
    algebra (attention_control.py:97-121) on the same accumulated maps."""
    lib = _lib.load()
    h = model.unet.handle
    # fill the store through a real UNet call of two images (rows prompt-major: [s0 s1 t0 t1])
    tok, te = model.tokenizer, model.text_encoder
    prompts = [synth.CAT_PROMPTS[0], synth.CAT_PROMPTS[0], synth.CAT_PROMPTS[1], synth.CAT_PROMPTS[1]]
    ctx = torch.cat([te(tok([""] * 4).input_ids)[0], te(tok(prompts).input_ids)[0]]).cuda().float().contiguous()
    _lib.check(lib.pnp_store_reset(h, G.stream()))
    ctrl = _lib.new_ctrl()
    for r, slot in ((4, 0), (6, 1), (5, 2), (7, 3)):
        ctrl.store_slot[r] = slot

    class Ctl:
        def descriptor(self, b):
            return ctrl

        def after_unet_call(self):
            pass

    model.unet.set_controller(Ctl())
    lat = torch.cat([synth.synth_latent(i) for i in range(4)]).cuda()
    model.unet(torch.cat([lat] * 2).contiguous(), 500, encoder_hidden_states=ctx)
    model.unet.set_controller(None)
    store = torch.empty(5, 2 * _lib.PNP_MAX_SLOTS, 8, 256, 77, device=cuda)
    _lib.check(lib.pnp_store_read(h, G.ptr(store), store.numel(), G.stream()))
    descs = (_lib.BlendDesc * 2)()
    words = [[2], [5]]  # token index 2 = "cat" in the source prompt; 5 = "cat" in the target prompt
    subs = [[], [4]]

    def norm_map(i, pr, word, pool):  # the quantity LocalBlend thresholds, for choosing thresholds that split the image
        mm = store[:, 2 * i + pr].cpu().double().reshape(40, 16, 16, 77)[..., word].mean(0)[None, None]
        if pool:
            mm = F.max_pool2d(mm, (3, 3), (1, 1), padding=(1, 1))
        return mm / mm.max()

    ths = []
    for i in range(2):
        vp = norm_map(i, 0, words[0][0], True).flatten().sort()[0].unique()
        vs = norm_map(i, 0, 4, False).flatten().sort()[0].unique()
        th_pool = float((vp[len(vp) // 2 - 1] + vp[len(vp) // 2]) / 2)   # between two distinct values: no borderline cell
        th_sub = float((vs[len(vs) * 4 // 5 - 1] + vs[len(vs) * 4 // 5]) / 2)
        ths.append((th_pool, th_sub))
    for i in range(2):
        d = descs[i]
        d.src_row, d.tgt_row, d.src_slot, d.tgt_slot = i, 2 + i, 2 * i, 2 * i + 1
        d.th_pool, d.th_sub = ths[i]
        d.nwords[0] = d.nwords[1] = 1
        d.words[0][0], d.words[1][0] = words[0][0], words[1][0]
        d.alpha[0][0] = d.alpha[1][0] = 1.0
        if subs[i]:
            d.nsub[0] = d.nsub[1] = 1
            d.sub_words[0][0] = d.sub_words[1][0] = subs[i][0]
            d.sub_alpha[0][0] = d.sub_alpha[1][0] = 1.0
    x = torch.randn(4, 4, 64, 64, generator=torch.Generator().manual_seed(4)).cuda()
    got = x.clone()
    masks = torch.zeros(2, 2, 4096, device=cuda)
    _lib.check(lib.pnp_local_blend_batch(h, G.ptr(got), 4, descs, 2, G.ptr(masks), G.stream()))
    torch.cuda.synchronize()
    for i in range(2):
        maps = store[:, 2 * i:2 * i + 2].cpu().double().permute(1, 0, 2, 3, 4).reshape(2, 40, 1, 16, 16, 77)

        def get_mask(word_src, word_tgt, pool, th):
            alpha = torch.zeros(2, 1, 1, 1, 1, 77, dtype=torch.float64)
            alpha[0, ..., word_src] = 1
            alpha[1, ..., word_tgt] = 1
            m = (maps * alpha).sum(-1).mean(1)
            if pool:
                m = F.max_pool2d(m, (3, 3), (1, 1), padding=(1, 1))
            mk = F.interpolate(m, size=(64, 64))
            mk = mk / mk.max(2, keepdim=True)[0].max(3, keepdim=True)[0]
            mk = mk.gt(th)
            return mk[:1] + mk

        import ctypes
        th_pool, th_sub = float(ctypes.c_float(ths[i][0]).value), float(ctypes.c_float(ths[i][1]).value)
        mask = get_mask(words[0][0], words[1][0], True, th_pool)
        if subs[i]:
            mask = mask * ~get_mask(subs[i][0], subs[i][0], False, th_sub)
        mask = mask.float()
        xc = x[[i, 2 + i]].cpu()
        ref = xc[:1] + mask * (xc - xc[:1])
        assert 0.0 < float(mask[0].mean()) < 1.0, (i, float(mask[0].mean()), float(mask[1].mean()))
        # cells whose normalised map sits within fp32 rounding of the threshold may fall on either side
        diff = (masks[i].cpu().reshape(2, 1, 64, 64) != mask).float().mean()
        assert float(diff) < 1e-2, (i, float(diff))
        agree = masks[i].cpu().reshape(2, 1, 64, 64) == mask
        assert torch.equal(torch.where(agree, got[[i, 2 + i]].cpu(), ref), ref), i


def test_config1_20_step_inversion_matches_the_reference(cuda):
    """BASELINE config 1: 20-step DDIM inversion of one latent, against the reference's own DirectInversion.ddim_loop on
    the vendored fp64 UNet (tests/golden/inversion_20steps.npz, oracle/make_golden.py pipeline_full)."""
    gold = os.path.join(os.path.dirname(__file__), "golden", "inversion_20steps.npz")
    if not os.path.exists(gold):
        pytest.fail("tests/golden/inversion_20steps.npz missing (python -m oracle.make_golden pipeline_full)")
    ref = torch.from_numpy(np.load(gold)["x_stars"])
    from pnpinversion_b200.batched import BatchedDirectInversionP2P, _encode_rows, _schedule, run_loop

    for table in ("float64", "float32"):
        m = FusedModel.synthetic(device="cuda:0", max_batch=4, table_dtype=table)
        ts, inv_t, inv_co, _ = _schedule(m, 20)
        assert ts == list(range(950, -1, -50))  # integer schedule, bit-exact
        ctx = _encode_rows(m, [synth.CAT_PROMPTS[0]]).contiguous()
        z = synth.synth_latent(0).cuda().contiguous()
        traj = torch.empty(21, 1, 4, 64, 64, device=cuda)
        run_loop(m, _lib.PNP_LOOP_INVERT, 20, 1, 1, inv_t, inv_co, 0.0, ctx, z, traj=traj)
        torch.cuda.synchronize()
        errs = [G.rel_l2(traj[k, 0].cpu(), ref[k]) for k in range(1, 21)]
        print(f"config 1 ({table} table): x_stars rel-L2 after 1/5/10/20 steps: "
              f"{errs[0]:.2e} {errs[4]:.2e} {errs[9]:.2e} {errs[19]:.2e}")
        assert torch.equal(traj[0, 0].cpu(), ref[0])
        assert max(errs) < 1e-2
        m.unet.close()


def test_skipping_the_reconstruction_pass_changes_nothing(model):
    """`minimal=True`: the reconstruction pass is not run, its source rows are taken from x_stars[0] (what the pass would
    return to 2e-5); inversion, offsets and the edit are bit-identical to the faithful run."""
    ed = P2PEditor(["directinversion+p2p"], "cuda:0", num_ddim_steps=4, model=model)
    src, tgt = synth.CAT_PROMPTS
    zs = torch.cat([synth.synth_latent(i) for i in range(2)]).cuda()
    calls0 = model.unet.kernel_launches()
    full = ed.edit_batch(zs, [src] * 2, [tgt] * 2, blend_word=BLEND, eq_params=EQ)
    calls1 = model.unet.kernel_launches()
    mini = ed.edit_batch(zs, [src] * 2, [tgt] * 2, blend_word=BLEND, eq_params=EQ, minimal=True)
    calls2 = model.unet.kernel_launches()
    torch.cuda.synchronize()
    assert torch.equal(full.x_stars, mini.x_stars) and torch.equal(full.noise_loss, mini.noise_loss)
    assert torch.equal(full.latents, mini.latents)
    assert (mini.reconstruct_latents[:2] - zs).abs().max() == 0 and (full.reconstruct_latents[:2] - zs).abs().max() < 2e-5
    assert (calls2 - calls1) < 0.8 * (calls1 - calls0)  # three 4-step loops instead of four
