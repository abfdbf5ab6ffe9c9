"""First contact of the round-2 kernels with the GPU: the 16-bit forward passes (bf16 / fp16 trunk, patch-GEMM head)
against their storage-point emulation (oracle/model_emul.py) and against the fp32 oracle.  Prints, asserts nothing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clasfv_b200
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
from oracle import fixtures, model_ref, model_emul
from oracle.model_emul import Config

CFG = {"bf16": Config(act="bf16", wtrunk="bf16", lateral="f16", head="patch", h1="bf16", h2="bf16", w2="bf16", wh="bf16", out="fp32"),
       "fp16": Config(act="f16", wtrunk="f16", lateral="f16", head="patch", h1="f16", h2="f16", w2="f16", wh="f16", out="fp32")}
sd = fixtures.calibrated_state_dict(0)
shapes = [((8, 32, 32), 2), ((16, 64, 48), 1), ((32, 112, 112), 1)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for shape, batch in shapes:
    x = fixtures.synthetic_clip(*shape, seed=13, batch=batch)
    seg_ref, mot_ref = model_ref.forward(sd, x)
    for prec in ("bf16", "fp16"):
        net = R2plus1D_18_MotionNet(pretrained=False, precision=prec); net.load_state_dict(sd); net = net.cuda().eval()
        t0 = time.time(); seg, mot = net(x.cuda()); torch.cuda.synchronize(); dt = time.time() - t0
        seg, mot = seg.float().cpu(), mot.float().cpu()
        seg_e, mot_e = model_emul.forward(sd, x, CFG[prec])
        w = shape[2]
        def m(a, b, am, bm):
            p, q = torch.softmax(a, 1), torch.softmax(b, 1)
            return (f"logit max|d| {float((a - b).abs().max()):.4f} (scale {float(b.std()):.2f}) softmax max {float((p - q).abs().max()):.5f} "
                    f"agree {float(((p[:, 1] > p[:, 0]) == (q[:, 1] > q[:, 0])).float().mean()) * 100:.4f}% flow max {float((am - bm).abs().max()) * w / 2:.5f} px "
                    f"mean {float((am - bm).abs().mean()) * w / 2:.5f} px")
        print(f"[{prec} {shape}x{batch}] finite {bool(torch.isfinite(seg).all() and torch.isfinite(mot).all())} ({dt:.2f}s)")
        print("    vs emulation :", m(seg, seg_e, mot, mot_e))
        print("    vs fp32 oracle:", m(seg, seg_ref, mot, mot_ref))
        print("    emulation vs fp32 oracle:", m(seg_e, seg_ref, mot_e, mot_ref), flush=True)
