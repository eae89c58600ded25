"""Per-class error budget of the fp16 pipeline (TEST / ANALYSIS TOOL, runs the CPU oracle; never part of the product).

The fused engine keeps every activation that crosses a kernel boundary in fp16 and feeds fp16 operands to the tensor
cores (fp32 accumulation).  This tool re-runs the oracle UNet (oracle/unet_ref.py) in fp64 with fp16 ROUNDING INSERTED at
the same places, one class of rounding points at a time, and reports the rel-L2 error each class contributes to one UNet
forward - the budget behind DESIGN.md section 2 (which roundings an fp32 residual stream removes, which need a hi/lo
split of the tensor-core operand).

    python tools/error_budget.py [--dtype float32|float64] [--classes a,b,...]  -> profiles/r2_error_budget.json
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import unet_ref  # noqa: E402
from pnpinversion_b200 import synth  # noqa: E402

CLASSES = ["stream", "inner_stream", "gn_out", "h1", "ln_out", "qkv", "probs", "attn_out", "geglu_out"]


def r16(x):
    return x.to(torch.float16).to(x.dtype)


def tf32(x):
    """round to nearest even on a 10-bit mantissa, fp32 exponent (cuDNN / cuBLAS TF32 operand conversion)"""
    b = x.to(torch.float32).contiguous().view(torch.int32)
    b = (b + 0x0FFF + ((b >> 13) & 1)) & ~0x1FFF
    return b.view(torch.float32).to(x.dtype)


def split16(x):
    """hi + lo fp16 pair (what a 2-MMA split-operand GEMM would consume): ~22 bits."""
    hi = x.to(torch.float16).to(x.dtype)
    lo = (x - hi).to(torch.float16).to(x.dtype)
    return hi + lo


class UNetEmu(unet_ref.UNetRef):
    """UNetRef with a rounding policy: self.rnd[class] in {None, 'fp16', 'split'}."""

    def __init__(self, sd, dtype, rnd):
        super().__init__(sd, dtype)
        self.rnd = rnd

    def q(self, x, cls):
        mode = self.rnd.get(cls)
        if mode == "fp16":
            return r16(x)
        if mode == "split":
            return split16(x)
        return x

    def resnet(self, x, temb, name):
        h = self.q(F.silu(self._gn(x, name + ".norm1", 1e-5)), "gn_out")
        h = self._conv(h, name + ".conv1")
        h = h + self._lin(F.silu(temb), name + ".time_emb_proj")[:, :, None, None]
        h = self.q(h, "h1")
        h = self.q(F.silu(self._gn(h, name + ".norm2", 1e-5)), "gn_out")
        h = self._conv(h, name + ".conv2")
        if name + ".conv_shortcut.weight" in self.p:
            x = self._conv(x, name + ".conv_shortcut", padding=0)
        return self.q(x + h, "stream")

    def attention(self, x, ctx, name, hook, place):
        is_cross = ctx is not None
        c = ctx if is_cross else x
        q = self.q(self._lin(x, name + ".to_q", bias=False), "qkv")
        k = self.q(self._lin(c, name + ".to_k", bias=False), "qkv")
        v = self.q(self._lin(c, name + ".to_v", bias=False), "qkv")

        def split(t):
            b, n, d = t.shape
            return t.reshape(b, n, unet_ref.HEADS, d // unet_ref.HEADS).permute(0, 2, 1, 3).reshape(b * unet_ref.HEADS, n, d // unet_ref.HEADS)

        q, k, v = split(q), split(k), split(v)
        scale = q.shape[-1] ** -0.5
        n = q.shape[1]
        out = torch.empty_like(q)
        for i0 in range(0, n, 512):
            sim = torch.einsum("bid,bjd->bij", q[:, i0:i0 + 512], k) * scale
            p = sim.softmax(dim=-1)
            if self.rnd.get("probs"):
                # the kernels round the UNNORMALISED exponentials (max-subtracted, in [0,1]) and normalise in fp32
                m = sim.max(dim=-1, keepdim=True)[0]
                e = self.q(torch.exp(sim - m), "probs")
                p = e / torch.exp(sim - m).sum(-1, keepdim=True)
            out[:, i0:i0 + 512] = torch.einsum("bij,bjd->bid", p, v)
        b8, n, d = out.shape
        out = out.reshape(b8 // 8, 8, n, d).permute(0, 2, 1, 3).reshape(b8 // 8, n, d * 8)
        return self._lin(self.q(out, "attn_out"), name + ".to_out.0")

    def transformer(self, x, ctx, name, hook, place):
        b, c, h, w = x.shape
        x_in = x
        x = self.q(self._gn(x, name + ".norm", 1e-6), "gn_out")
        x = self.q(self._conv(x, name + ".proj_in", padding=0), "inner_stream")
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        t = name + ".transformer_blocks.0"
        x = self.q(self.attention(self.q(self._ln(x, t + ".norm1"), "ln_out"), None, t + ".attn1", hook, place) + x, "inner_stream")
        x = self.q(self.attention(self.q(self._ln(x, t + ".norm2"), "ln_out"), ctx, t + ".attn2", hook, place) + x, "inner_stream")
        y = self._lin(self.q(self._ln(x, t + ".norm3"), "ln_out"), t + ".ff.net.0.proj")
        val, gate = y.chunk(2, dim=-1)
        x = self.q(self._lin(self.q(val * F.gelu(gate), "geglu_out"), t + ".ff.net.2") + x, "inner_stream")
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return self.q(self._conv(x, name + ".proj_out", padding=0) + x_in, "stream")

    def _conv(self, x, name, stride=1, padding=1):
        if self.rnd.get("tf32_conv"):
            # what the reference's own GPU path does by default: torch.backends.cudnn.allow_tf32 is True, so every cuDNN
            # convolution rounds its operands to TF32 (10 mantissa bits, the width of fp16; fp32 exponent range).  The
            # synthetic weights are fp16 values and therefore exact in TF32: only the activation operand is rounded.
            x = tf32(x)
        out = super()._conv(x, name, stride, padding)
        if name == "conv_in" or "samplers" in name:
            out = self.q(out, "stream")
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--configs", default="")
    ap.add_argument("--out", default="profiles/r2_error_budget.json")
    args = ap.parse_args()
    if args.threads:
        torch.set_num_threads(args.threads)
    dt = getattr(torch, args.dtype)
    sd = synth.synth_unet_state_dict(0)
    tok, te = synth.FakeTokenizer(), synth.SynthTextEncoder()
    ctx = te(tok([synth.CAT_PROMPTS[0]]).input_ids)[0]
    x = synth.synth_latent(0)
    t = 981
    configs = {"exact": {}}
    for c in CLASSES:
        configs["only_" + c] = {c: "fp16"}
    configs["all_fp16 (the engine today)"] = {c: "fp16" for c in CLASSES}
    configs["fp32_streams (stream+inner_stream+h1 exact)"] = {c: "fp16" for c in CLASSES if c not in ("stream", "inner_stream", "h1")}
    configs["fp32_streams + split gn_out/ln_out"] = {c: ("split" if c in ("gn_out", "ln_out") else "fp16")
                                                     for c in CLASSES if c not in ("stream", "inner_stream", "h1")}
    configs["fp32_streams + split all GEMM A operands (gn_out, ln_out, attn_out, geglu_out)"] = {
        c: ("split" if c in ("gn_out", "ln_out", "attn_out", "geglu_out") else "fp16")
        for c in CLASSES if c not in ("stream", "inner_stream", "h1")}
    configs["fp32_streams + split all GEMM A operands + fp32 probs"] = {
        c: ("split" if c in ("gn_out", "ln_out", "attn_out", "geglu_out") else "fp16")
        for c in CLASSES if c not in ("stream", "inner_stream", "h1", "probs")}
    configs["everything split: fp32 streams, hi/lo GEMM and attention operands (gn_out, ln_out, attn_out, geglu_out, qkv), fp32 probs"] = {
        c: "split" for c in ("gn_out", "ln_out", "attn_out", "geglu_out", "qkv")}
    configs["REFERENCE on a GPU with its defaults: exact fp32 except cuDNN TF32 convolutions (torch.backends.cudnn.allow_tf32 = True)"] = {
        "tf32_conv": True}
    if args.configs:
        want = args.configs.split(",")
        configs = {k: v for k, v in configs.items() if k == "exact" or any(w in k for w in want)}
    res, ref = {}, None
    outp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), args.out)
    if args.configs and os.path.exists(outp):
        res = json.load(open(outp)).get("results", {})
    with torch.no_grad():
        for name, rnd in configs.items():
            t0 = time.time()
            out = UNetEmu(sd, dt, rnd)(x, t, ctx)
            if ref is None:
                ref = out
                continue
            e = float((out.double() - ref.double()).norm() / ref.double().norm())
            res[name] = e
            print(f"{name:90s} rel-L2 {e:.3e}   ({time.time() - t0:.0f}s)", flush=True)
            with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), args.out), "w") as f:
                json.dump({"what": "rel-L2 of one B=1 UNet forward (t=981, cat prompt, synthetic weights) vs the same "
                                   "oracle without rounding, fp16 rounding inserted per class of rounding points",
                           "dtype": args.dtype, "results": res}, f, indent=1)


if __name__ == "__main__":
    main()
