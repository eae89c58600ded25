"""Where along the down path do two batch sizes of the same row start to differ?  Reads the 12 skip tensors
(pnp_debug_read) after a B=1 and a B=2 / B=4 forward of identical rows and prints rel-L2 of row 0 per tensor."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnpinversion_b200 import _lib, synth
from pnpinversion_b200.model import FusedModel

m = FusedModel.synthetic(max_batch=4)
lib = _lib.load()
tok, te = m.tokenizer, m.text_encoder
ctx1 = te(tok([synth.CAT_PROMPTS[0]]).input_ids)[0].cuda().float()
x1 = synth.synth_latent(0).cuda()
C_ = [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
HW = [64, 64, 64, 32, 32, 32, 16, 16, 16, 8, 8, 8]
names = ["conv_in", "down0.res0+xf0", "down0.res1+xf1", "down0.downsample", "down1.res0+xf", "down1.res1+xf", "down1.down",
         "down2.res0+xf", "down2.res1+xf", "down2.down", "down3.res0", "down3.res1"]

def run(B):
    x = x1.expand(B, -1, -1, -1).contiguous(); ctx = ctx1.expand(B, -1, -1).contiguous()
    out = m.unet(x, 981, encoder_hidden_states=ctx)["sample"]
    torch.cuda.synchronize()
    sk = []
    for i in range(12):
        n = B * HW[i] * HW[i] * C_[i]
        buf = torch.empty(n, dtype=torch.float16, device="cuda")
        got = C.c_int64()
        _lib.check(lib.pnp_debug_read(m.unet.handle, B, i, C.c_void_p(buf.data_ptr()), n, C.byref(got), _lib.current_stream_ptr()))
        torch.cuda.synchronize()
        sk.append(buf[: HW[i] * HW[i] * C_[i]].float().cpu())  # row 0
    return out[:1].cpu(), sk

rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
o1, s1 = run(1)
for B in (2, 4):
    oB, sB = run(B)
    print(f"--- B={B} vs B=1 (row 0): eps {rel(oB, o1):.3e}")
    for i in range(12):
        nd = int((sB[i] != s1[i]).sum())
        print(f"  skip {i:2d} {names[i]:18s} rel-L2 {rel(sB[i], s1[i]):.3e}   differing elements {nd} / {s1[i].numel()}")
