"""TEST INFRASTRUCTURE ONLY -- CPU restatement (fp64) of the reference's EDICT coupled loop,
`models/edict/edict_functions.py::coupled_stablediffusion` (:707-956) with forward_step / reverse_step (:621-684),
the mixing layers (:854-859, :931-936), the leapfrog order (:862-880) and the attention reuse of its P2P mode
(:250-297: attn1 replaced wholesale, attn2 = P*(1-mask) + P_saved[..., indices]*mask), driving oracle/unet_ref.py.

PARITY PINNED: tests/golden/edict_2steps.npz holds outputs of the reference's own `coupled_stablediffusion`, run
UNMODIFIED in the build container (oracle/ref_shim.load_reference_edict compiles the file's function definitions and
supplies the module globals - the vendored fp64 UNet with the synthetic weights, a CLIP stand-in, the whitespace
tokenizer - because the module itself downloads its models from the hub and moves them to 'cuda' at import time, :36-53;
generator: `python -m oracle.make_golden edict`).  tests/test_oracle_cpu.py replays the recorded UNet calls through
`coupled` (same call order, inputs and resulting latents to 1e-6) and checks the attention save / reuse hook on a full-size
forward pair against the recorded predictions (1e-6).  Only tests/ may import this module.
"""
from __future__ import annotations

import torch


def _alpha(ac, final, t):
    return ac[t] if t >= 0 else final


def forward_step(ac, final, eps, t, x, ratio):
    a_t, a_p = _alpha(ac, final, t), _alpha(ac, final, t - ratio)
    q = (a_t / a_p) ** 0.5
    return (1.0 / q) * x - (1.0 / q) * ((1 - a_t) ** 0.5) * eps + ((1 - a_p) ** 0.5) * eps


def reverse_step(ac, final, eps, t, x, ratio):
    a_t, a_p = _alpha(ac, final, t), _alpha(ac, final, t - ratio)
    q = (a_t / a_p) ** 0.5
    return q * x + ((1 - a_t) ** 0.5) * eps - q * ((1 - a_p) ** 0.5) * eps


class _Reuse:
    """The save/use flags of edict_functions.py:250-327 as an attention hook for two consecutive UNet calls."""

    def __init__(self, mask, indices):
        self.mask, self.indices = mask, indices
        self.saved = []
        self.mode = "off"
        self.k = 0

    def __call__(self, attn, is_cross, place):
        if self.mode == "save":
            self.saved.append(attn.clone())
        elif self.mode == "use":
            last = self.saved[self.k]
            self.k += 1
            if is_cross:
                attn = attn * (1 - self.mask) + last[..., self.indices] * self.mask
            else:
                attn = last
        return attn


def coupled(unet, ac, final, timesteps_all, pair, emb_u, emb_c, emb_e=None, mask=None, indices=None, guidance=7.0,
            steps=50, t_limit=0, reverse=False, mix=0.93):
    ratio = 1000 // steps
    ts = timesteps_all[t_limit:]
    if reverse:
        ts = ts.flip(0)
    pair = [p.clone() for p in pair]
    n = len(ts)
    for i, t in enumerate(ts):
        t = int(t)
        if reverse:
            new = [l.clone() for l in pair]
            new[1] = (new[1] - (1 - mix) * new[0]) / mix
            new[0] = (new[0] - (1 - mix) * new[1]) / mix
            pair = new
        for k in range(2):
            li = (k + ((n - (i + 1)) + 1) % 2) % 2 if reverse else (k + i % 2) % 2
            lj = (li + 1) % 2
            x = pair[lj]
            e_u = unet(x, t, emb_u, None)
            if emb_e is not None:
                hook = _Reuse(mask, indices)
                hook.mode = "save"
                e_c = unet(x, t, emb_c, hook)
                hook.mode = "use"
                e_c = unet(x, t, emb_e, hook)
            else:
                e_c = unet(x, t, emb_c, None)
            e = e_u + guidance * (e_c - e_u)
            step = reverse_step if reverse else forward_step
            pair[li] = step(ac, final, e, t, pair[li], ratio)
        if not reverse:
            new = [l.clone() for l in pair]
            new[0] = mix * new[0] + (1 - mix) * new[1]
            new[1] = (1 - mix) * new[0] + mix * new[1]
            pair = new
    return pair
