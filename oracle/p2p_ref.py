"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's DDIM-inversion + PnP latent-rectification loops and
of the Prompt-to-Prompt controller algebra, on materialised attention probabilities (the way the reference does it).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module.  It follows:
  * schedule / steps .... models/p2p/inversion.py:247-270 (prev_step / next_step), models/p2p/scheduler_dev.py:40-51,91-94,
                          table my_diffusers/schedulers/scheduling_ddim.py:105,112-119,152-154
  * loops ............... models/p2p/inversion.py:308-319 (ddim_loop), :375-391 (offset_calculate),
                          models/p2p/p2p_guidance_forward.py:103-116,135-173
  * controllers ......... models/p2p/attention_control.py:95-147 (LocalBlend), :178-190, :214-248 (AttentionStore),
                          :258-282 (AttentionControlEdit), :317-363 (Refine / Reweight)
Pinned by tests/test_oracle_cpu.py against fixtures produced by the reference's own code (oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

NUM_TRAIN = 1000


def alphas_cumprod(table: str = "float32") -> torch.Tensor:
    """'float32': diffusers>=0.10 builds betas with torch.linspace(dtype=float32) (the P2P/MasaCtrl paths);
    'float64': the vendored 0.3.0 scheduler builds it in numpy float64 (scheduling_ddim.py:105,112-113)."""
    if table == "float32":
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, NUM_TRAIN, dtype=torch.float32) ** 2
        return torch.cumprod(1.0 - betas, dim=0)
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, NUM_TRAIN, dtype=np.float64) ** 2
    return torch.from_numpy(np.cumprod(1.0 - betas, axis=0))


def timesteps(n: int) -> torch.Tensor:
    """scheduling_ddim.py:152-154: arange(0, 1000, 1000 // n)[::-1], int64."""
    return torch.from_numpy(np.arange(0, NUM_TRAIN, NUM_TRAIN // n)[::-1].copy()).to(torch.int64)


class Schedule:
    def __init__(self, n: int, table: str = "float32"):
        self.n = n
        self.ac = alphas_cumprod(table)
        self.final = self.ac[0]  # set_alpha_to_one=False, p2p_editor.py:22
        self.timesteps = timesteps(n)
        self.ratio = NUM_TRAIN // n

    def prev_step(self, eps, t: int, x):
        """inversion.py:247-255."""
        prev_t = t - self.ratio
        a_t = self.ac[t]
        a_prev = self.ac[prev_t] if prev_t >= 0 else self.final
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps

    def next_step(self, eps, t: int, x):
        """inversion.py:262-270."""
        cur_t, next_t = min(t - self.ratio, 999), t
        a_t = self.ac[cur_t] if cur_t >= 0 else self.final
        a_next = self.ac[next_t]
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_next ** 0.5 * x0 + (1 - a_next) ** 0.5 * eps


# ---------------------------------------------------------------------------------------------------- controllers
class StoreController:
    """AttentionStore (attention_control.py:214-248): acts on the cond half only (:184), keeps maps with <= 32^2
    queries, sums them over steps.  Note the aliasing (:224 stores a view that a later edit writes through)."""

    def __init__(self):
        self.num_att_layers = 32
        self.reset()

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0
        self.step_store = self._empty()
        self.attention_store = {}

    @staticmethod
    def _empty():
        return {f"{p}_{k}": [] for p in ("down", "mid", "up") for k in ("cross", "self")}

    def forward(self, attn, is_cross, place):
        if attn.shape[1] <= 32 ** 2:
            self.step_store[f"{place}_{'cross' if is_cross else 'self'}"].append(attn)
        return attn

    def __call__(self, attn, is_cross, place):
        h = attn.shape[0]
        attn[h // 2:] = self.forward(attn[h // 2:], is_cross, place)
        self._layer_done()
        return attn

    def passthrough(self, is_cross, place):
        """Bookkeeping of `__call__` for a layer whose maps this controller neither stores nor edits (self-attention
        with more than 32^2 queries): lets the UNet restatement skip materialising the map."""
        assert not is_cross
        self._layer_done()

    def _layer_done(self):
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            self.between_steps()

    def between_steps(self):
        if not self.attention_store:
            self.attention_store = self.step_store
        else:
            for k in self.attention_store:
                for i in range(len(self.attention_store[k])):
                    self.attention_store[k][i] += self.step_store[k][i]
        self.step_store = self._empty()

    def step_callback(self, x_t):
        return x_t


class EditController(StoreController):
    """AttentionRefine optionally wrapped by AttentionReweight, with LocalBlend (attention_control.py:251-363),
    for one (source, target) prompt pair.  Tables are given explicitly (they come from the host-side aligner)."""

    def __init__(self, num_steps, cross_replace_alpha, self_replace_steps, mapper, alphas, equalizer=None,
                 blend_alpha_layers=None, blend_start=0.2, blend_th=0.3):
        super().__init__()
        self.batch_size = 2
        self.cross_replace_alpha = cross_replace_alpha  # (num_steps+1, 1, 1, 1, 77)
        self.num_self_replace = (0, int(num_steps * self_replace_steps))
        self.mapper = mapper.long()  # (1,77)
        self.alphas = alphas.reshape(1, 1, 1, -1)
        self.equalizer = equalizer  # (1,77) or None
        self.blend = blend_alpha_layers  # (2,1,1,1,1,77) or None
        self.start_blend = int(blend_start * num_steps)
        self.th = blend_th
        self.counter = 0

    def forward(self, attn, is_cross, place):
        super().forward(attn, is_cross, place)
        if is_cross or (self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1]):
            h = attn.shape[0] // self.batch_size
            attn = attn.reshape(self.batch_size, h, *attn.shape[1:])
            base, repl = attn[0], attn[1:]
            if is_cross:
                a = self.cross_replace_alpha[self.cur_step]
                new = base[:, :, self.mapper].permute(2, 0, 1, 3) * self.alphas + repl * (1 - self.alphas)
                if self.equalizer is not None:
                    new = new * self.equalizer[:, None, None, :]
                attn[1:] = new * a + (1 - a) * repl
            else:
                if repl.shape[2] <= 32 ** 2:
                    attn[1:] = base.unsqueeze(0).expand(repl.shape[0], *base.shape)
            attn = attn.reshape(self.batch_size * h, *attn.shape[2:])
        return attn

    def step_callback(self, x_t):
        if self.blend is None:
            return x_t
        self.counter += 1
        if self.counter > self.start_blend:
            maps = self.attention_store["down_cross"][2:4] + self.attention_store["up_cross"][:3]
            maps = [m.reshape(2, -1, 1, 16, 16, 77) for m in maps]
            maps = torch.cat(maps, dim=1)
            m = (maps * self.blend).sum(-1).mean(1)
            m = F.max_pool2d(m, (3, 3), (1, 1), padding=(1, 1))
            mask = F.interpolate(m, size=(64, 64))
            mask = mask / mask.max(2, keepdim=True)[0].max(3, keepdim=True)[0]
            mask = mask.gt(self.th)
            mask = mask[:1] + mask
            x_t = x_t[:1] + mask.to(x_t.dtype) * (x_t - x_t[:1])
        return x_t


# ---------------------------------------------------------------------------------------------------- loops
def ddim_loop(unet: Callable, sched: Schedule, z0, cond_ctx):
    """inversion.py:308-319: x_stars[0] = z0, 50 x { eps = unet(x, t, cond_src) ; x = next_step }."""
    xs = [z0]
    x = z0.clone()
    for i in range(sched.n):
        t = int(sched.timesteps[sched.n - i - 1])
        eps = unet(x, t, cond_ctx, None)
        x = sched.next_step(eps, t, x)
        xs.append(x)
    return xs


def offset_calculate(unet: Callable, sched: Schedule, x_stars: List[torch.Tensor], context, guidance: float):
    """inversion.py:375-391: both prompt rows follow the source latents; loss = x_stars[...] - rec."""
    n_prompts = context.shape[0] // 2
    cur = torch.cat([x_stars[-1]] * n_prompts)
    losses = []
    for i in range(sched.n):
        prev = torch.cat([x_stars[len(x_stars) - i - 2]] * cur.shape[0])
        t = int(sched.timesteps[i])
        eps = unet(torch.cat([cur] * 2), t, context, None)
        eu, ec = eps.chunk(2)
        e = eu + guidance * (ec - eu)
        rec = sched.prev_step(e, t, cur)
        loss = prev - rec
        losses.append(loss)
        cur = rec + loss
    return losses


def guidance_forward(unet: Callable, sched: Schedule, x_T, context, guidance: float, noise_loss, controller,
                     add_offset: bool = True):
    """p2p_guidance_forward.py:103-116,135-173 with the rectification at :113-114."""
    n_prompts = context.shape[0] // 2
    lat = x_T.expand(n_prompts, *x_T.shape[1:]).clone()
    for i in range(sched.n):
        t = int(sched.timesteps[i])
        eps = unet(torch.cat([lat] * 2), t, context, controller)
        eu, ec = eps.chunk(2)
        e = eu + guidance * (ec - eu)
        lat = sched.prev_step(e, t, lat)
        if add_offset:
            lat = torch.cat((lat[:1] + noise_loss[i][:1], lat[1:]))
        if controller is not None:
            lat = controller.step_callback(lat)
    return lat
