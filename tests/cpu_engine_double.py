"""TEST DOUBLE (CPU): stands in for the CUDA engine so that the product's HOST loops (inversion.py,
p2p_guidance_forward.py: timestep / index arithmetic, which latent feeds which step, where the rectification goes) can run
on a machine without a GPU and be compared with the reference's loops over a whole 50-step schedule.  It is never used by
the product: tests monkeypatch `fused_step` in the product modules with `cpu_fused_step` and pass a `FakeUNet`.

`cpu_fused_step` evaluates what `csrc/epilogue.cu::step_epilogue_kernel` evaluates (pnp_attn.h: StepParams), in the same
order, with fp32 torch ops."""
import torch


def cpu_fused_step(engine, x, eps_c, coeffs, eps_u=None, guidance=1.0, target=None, loss_out=None, noise_loss=None,
                   add_mask=0, out=None, loss_scale=1.0):
    c0, c1, c2, c3 = coeffs
    eps = eps_c if eps_u is None else eps_u + guidance * (eps_c - eps_u)
    x_new = c2 * ((x - c1 * eps) / c0) + c3 * eps
    if target is not None:  # OFFSET: loss = target - x_new ; x_new = x_new + loss (row r uses target row r % target_rows)
        tgt = target[[r % target.shape[0] for r in range(x.shape[0])]]
        loss = tgt - x_new
        if loss_scale != 1.0:
            loss = loss * loss_scale
        if loss_out is not None:
            loss_out.copy_(loss)
        x_new = x_new + loss
    if noise_loss is not None:  # RECTIFY: rows flagged in add_mask
        rows = [r for r in range(x.shape[0]) if (add_mask >> r) & 1]
        x_new = x_new.clone()
        x_new[rows] = x_new[rows] + noise_loss[rows]
    if out is not None:
        out.copy_(x_new)
        return out
    return x_new


class FakeUNet:
    """A cheap deterministic, nonlinear stand-in with the UNet's call surface: the prediction depends on the latent, the
    timestep and (per batch row) the text embedding, so a loop that feeds the wrong latent / timestep / context row to a
    step diverges immediately."""

    in_channels = 4
    handle = None

    def __init__(self):
        self.calls = []
        self._controller = None

    def set_controller(self, controller):
        self._controller = controller

    def named_children(self):  # the reference's register_attention_control walks the module tree: nothing to patch here
        return []

    def __call__(self, x, t, encoder_hidden_states=None):
        self.calls.append((tuple(x.shape), int(t)))
        s = encoder_hidden_states.float().mean(dim=(1, 2)).reshape(-1, 1, 1, 1)
        eps = torch.tanh(0.7 * x.float() + 0.3 * torch.roll(x.float(), 1, dims=-1) + 40.0 * s + float(t) / 1000.0) * 0.8
        return {"sample": eps}


class FakeLib:
    """The two C entry points the EDICT host loop calls directly (pnp_edict_mix), on CPU tensors addressed by pointer."""

    @staticmethod
    def _view(ptr, numel):
        import ctypes as C

        addr = ptr.value if hasattr(ptr, "value") else int(ptr)
        return torch.frombuffer((C.c_float * numel).from_address(addr), dtype=torch.float32)

    def pnp_edict_mix(self, handle, p0, p1, n, w, reverse, stream):
        """csrc/epilogue.cu::edict_mix_kernel: the mixing layers of edict_functions.py:854-859 (reverse) / :931-936."""
        numel = int(n) * 4 * 64 * 64
        x, y = self._view(p0, numel), self._view(p1, numel)
        w = float(w)
        if reverse:
            y.copy_((y - (1 - w) * x) / w)
            x.copy_((x - (1 - w) * y) / w)
        else:
            x.copy_(w * x + (1 - w) * y)
            y.copy_((1 - w) * x + w * y)
        return 0
