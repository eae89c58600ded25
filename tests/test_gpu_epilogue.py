"""Fused step epilogue (CFG + DDIM step / inverse step + offset + rectification) and LocalBlend vs the CPU oracle
(oracle/p2p_ref.py, which restates models/p2p/inversion.py:247-270,383-389 and attention_control.py:97-121)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import p2p_ref
from pnpinversion_b200 import _lib
from pnpinversion_b200.scheduler import DDIMSchedulerDev, fused_step, step_coefficients
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(cuda):
    h = C.c_void_p()
    _lib.check(_lib.load().pnp_create(0, 4, C.byref(h)))
    yield h
    _lib.load().pnp_destroy(h)


@pytest.mark.parametrize("table", ["float32", "float64"])
def test_prev_and_next_step_match_oracle(engine, cuda, table):
    sch = DDIMSchedulerDev(engine=engine, table_dtype=table)
    sch.set_timesteps(50)
    ora = p2p_ref.Schedule(50, table)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 64, 64, generator=g)
    e = torch.randn(2, 4, 64, 64, generator=g)
    for t in (980, 500, 20, 0):
        got = sch.step(e.to(cuda), t, x.to(cuda))["prev_sample"].cpu()
        ref = ora.prev_step(e.to(ora.ac.dtype), t, x.to(ora.ac.dtype))
        assert G.rel_l2(got, ref) < 1e-6, (t, table)
        co = step_coefficients(sch.alphas_cumprod, sch.final_alpha_cumprod, min(t - 20, 999), t)
        got = fused_step(engine, x.to(cuda), e.to(cuda), co).cpu()
        ref = ora.next_step(e.to(ora.ac.dtype), t, x.to(ora.ac.dtype))
        assert G.rel_l2(got, ref) < 1e-6, (t, table)
    # identity steps at t = 0 (SURVEY.md section 4 invariant b)
    co = step_coefficients(sch.alphas_cumprod, sch.final_alpha_cumprod, -20, 0)
    assert G.rel_l2(fused_step(engine, x.to(cuda), e.to(cuda), co).cpu(), x) < 1e-6


def test_fp32_path_is_bit_exact_with_eager_fp32(engine, cuda):
    """With the float32 table the kernel reproduces the reference's eager fp32 expression sequence bit for bit."""
    sch = DDIMSchedulerDev(engine=engine, table_dtype="float32")
    sch.set_timesteps(50)
    ac = sch.alphas_cumprod
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 64, 64, generator=g)
    eu = torch.randn(2, 4, 64, 64, generator=g)
    ec = torch.randn(2, 4, 64, 64, generator=g)
    nl = torch.randn(2, 4, 64, 64, generator=g) * 0.01
    t, pt = 500, 480
    eps = eu + 7.5 * (ec - eu)
    a_t, a_p = ac[t], ac[pt]
    x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    prev = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps
    ref = torch.cat((prev[:1] + nl[:1], prev[1:]))
    co = step_coefficients(ac, sch.final_alpha_cumprod, t, pt)
    got = fused_step(engine, x.to(cuda), ec.to(cuda), co, eps_u=eu.to(cuda), guidance=7.5, noise_loss=nl.to(cuda),
                     add_mask=1).cpu()
    assert torch.equal(got, ref)


def test_offset_mode(engine, cuda):
    sch = DDIMSchedulerDev(engine=engine, table_dtype="float32")
    sch.set_timesteps(50)
    ora = p2p_ref.Schedule(50, "float32")
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 64, 64, generator=g)
    eu, ec = torch.randn(2, 4, 64, 64, generator=g), torch.randn(2, 4, 64, 64, generator=g)
    target = torch.randn(1, 4, 64, 64, generator=g)
    t = 740
    rec = ora.prev_step(eu + 7.5 * (ec - eu), t, x)
    loss_ref = torch.cat([target] * 2) - rec
    cur_ref = rec + loss_ref
    co = step_coefficients(sch.alphas_cumprod, sch.final_alpha_cumprod, t, t - 20)
    loss = torch.empty(2, 4, 64, 64, device=cuda)
    cur = fused_step(engine, x.to(cuda), ec.to(cuda), co, eps_u=eu.to(cuda), guidance=7.5, target=target.to(cuda),
                     loss_out=loss)
    assert G.rel_l2(loss.cpu(), loss_ref) < 1e-6 and G.rel_l2(cur.cpu(), cur_ref) < 1e-6
    # the branch lands on the target up to one rounding (rec + (target - rec))
    assert (cur.cpu() - target).abs().max() < 1e-5


def test_local_blend_kernel_vs_oracle_via_pipeline_store(cuda):
    """Drive one real UNet call with store slots on, read the store back, run the LocalBlend kernel and compare with the
    oracle's LocalBlend applied to the very same maps."""
    from pnpinversion_b200 import synth
    from pnpinversion_b200.attention_control import make_controller, register_attention_control
    from pnpinversion_b200.model import FusedModel

    model = FusedModel.synthetic(device="cuda:0", max_batch=4)
    prompts = list(synth.CAT_PROMPTS)
    tok, te = model.tokenizer, model.text_encoder
    ctx = torch.cat([te(tok([""] * 2).input_ids)[0], te(tok(prompts).input_ids)[0]]).to(cuda, torch.float32)
    ctrl = make_controller(model, prompts, False, {"default_": 0.4}, 0.6, blend_words=(("cat",), ("cat",)),
                           num_ddim_steps=4)
    register_attention_control(model, ctrl)
    lat = torch.cat([synth.synth_latent(0), synth.synth_latent(1)]).to(cuda)
    model.unet(torch.cat([lat] * 2), 500, encoder_hidden_states=ctx)
    store = torch.empty(5, 2 * _lib.PNP_MAX_SLOTS, 8, 256, 77, device=cuda)
    _lib.check(_lib.load().pnp_store_read(model.unet.handle, C.c_void_p(store.data_ptr()), store.numel(),
                                          _lib.current_stream_ptr()))
    x = torch.randn(2, 4, 64, 64, generator=torch.Generator().manual_seed(4)).to(cuda)
    got = ctrl.step_callback(x.clone())  # counter 1 > start_blend 0 -> blends
    torch.cuda.synchronize()
    # oracle LocalBlend on the same maps
    maps = store[:, :2].cpu().double()  # (5,2,8,256,77)
    maps = maps.permute(1, 0, 2, 3, 4).reshape(2, 40, 1, 16, 16, 77)
    alpha = ctrl.local_blend.alpha_layers.double().reshape(2, 1, 1, 1, 1, 77)
    m = (maps * alpha).sum(-1).mean(1)
    m = F.max_pool2d(m, (3, 3), (1, 1), padding=(1, 1))
    mask = F.interpolate(m, size=(64, 64))
    mask = mask / mask.max(2, keepdim=True)[0].max(3, keepdim=True)[0]
    mask = mask.gt(0.3)
    mask = (mask[:1] + mask).float()
    xc = x.cpu()
    ref = xc[:1] + mask * (xc - xc[:1])
    frac = float(mask[1].mean())
    assert 0.0 < frac < 1.0, frac  # a non-trivial mask
    assert torch.equal(got.cpu(), ref)
    model.unet.close()
