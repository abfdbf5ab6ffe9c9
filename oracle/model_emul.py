"""Oracle: the CLAS-FV forward pass evaluated with the *storage points* of the CUDA bf16 path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``oracle/model_ref.py`` is the reference's arithmetic (fp32 everywhere).  This module is the same network
evaluated the way ``csrc/`` evaluates it in bf16 mode: every value is rounded exactly where the CUDA path
stores it, every sum is an fp32 accumulation of exact 16-bit x 16-bit products, BatchNorm is folded into the
weights (``csrc/api.cu:pack_conv_bn``) and ``comb_1`` is commuted with the up-sampling (``api.cu`` "decoder").
It exists to separate two things the fp32 oracle cannot tell apart ("next round" item 1a):

* *implementation error* - the CUDA path must agree with this emulation to within a couple of roundings of the
  stored type (it differs only by the order of the fp32 additions), and
* *storage noise* - how far a 16-bit evaluation of this network is from the fp32 reference on a given set of
  weights, whatever the implementation.

Storage points (``Config``), with the file that fixes each of them:

====================  =================================================================================
stem 1x7x7            fp32 weights (BN folded) on the fp32 clip, + shift, ReLU -> ``act``          conv_simt.cu:stem_conv_kernel
every trunk conv      ``act`` input x bf16(w * bn_scale), fp32 accumulate, + bn_shift (fp32),
                      + residual, ReLU -> ``act``                                                   conv_umma.cu epilogue
residual stream       block outputs are additionally kept in ``residual`` precision for the next
                      block's skip connection (``"act"`` = the same rounded tensor)                api.cu:run_block
lateral 1x1x1         ``act`` input x bf16(w1 * s1), fp32 accumulate -> ``lateral``                 api.cu:run_tail
temporal pre-pass     l0 * a + l1 * b in fp32 -> ``lateral``                                        decoder_umma.cu:temporal_upsample_kernel
head, interpolation   ``"row"``: R = wH0*row0 + wH1*row1 (fp32) -> fp16; W weights fp16; b1 hi+lo
                      ``"patch"``: weights wH*wW rounded to ``lateral``'s type, K = the patch's
                      low-resolution pixels, b1 hi+lo in that type                                  decoder_umma.cu
head, comb_2 / heads  relu -> bf16; x bf16(W2 * s2) + b2 (hi+lo bf16); relu -> bf16; x bf16(Wh)
                      + bh (fp32); softmax / tanh in fp32 -> ``out``                                decoder_umma.cu epilogues
====================  =================================================================================
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

from oracle.model_ref import BN_EPS, LAYERS, strip_module_prefix

_DT = {"bf16": torch.bfloat16, "f16": torch.float16, "fp32": torch.float32}


@dataclass(frozen=True)
class Config:
    act: str = "bf16"            # trunk activation storage
    residual: str = "act"        # "act" | "fp32": precision of the skip-connection operand
    lateral: str = "bf16"        # "bf16" | "f16": lateral maps g_l (and the interpolation weights of the patch head)
    head: str = "row"            # "row" (round-1 head) | "patch" (2-D patch GEMM head)
    out: str = "bf16"            # "bf16" | "fp32": prob / logits / motion storage
    split_act: tuple = ()        # names of convolutions whose input is fed as hi + lo (two MMAs per product)
    h1: str = "bf16"             # relu(comb_1) as the A operand of comb_2: "bf16" | "bf16x2" (hi + lo) | "tf32" | "fp32"
    h2: str = "bf16"             # relu(comb_2) as the A operand of the heads
    w2: str = "bf16"             # comb_2 weights: "bf16" | "bf16x2" | "tf32" | "fp32"
    wh: str = "bf16"             # head weights
    wtrunk: str = "bf16"         # trunk / lateral weights
    exact: tuple = ()            # ablation only: name prefixes of convolutions evaluated with unrounded weights and outputs


def rnd(x, kind):
    """Round to the storage type and come back to fp32 (fp16 saturates like cvt.rn.satfinite)."""
    if kind == "fp32":
        return x
    if kind == "bf16x2":                                   # hi + lo pair of bf16 values (16 significant bits)
        hi = x.to(torch.bfloat16).float()
        return hi + (x - hi).to(torch.bfloat16).float()
    if kind == "tf32":                                     # 10 explicit mantissa bits, round to nearest even
        i = x.contiguous().view(torch.int32)
        i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
        return i.view(torch.float32)
    if kind == "f16":
        return x.clamp(-65504.0, 65504.0).to(torch.float16).float()
    return x.to(torch.bfloat16).float()


def bn_affine(sd, key):
    s = sd[key + ".weight"] / torch.sqrt(sd[key + ".running_var"] + BN_EPS)
    return s, sd[key + ".bias"] - sd[key + ".running_mean"] * s


def _conv(x, sd, conv_key, bn_key, stride, pad, cfg, residual=None, relu=True, name=""):
    s, b = bn_affine(sd, bn_key)
    w = sd[conv_key + ".weight"] * s.view(-1, 1, 1, 1, 1)
    if not name.startswith(cfg.exact or ("\0",)):
        w = rnd(w, cfg.wtrunk)
    if name in cfg.split_act:
        hi = rnd(x, "bf16")
        y = F.conv3d(hi, w, None, stride, pad) + F.conv3d(rnd(x - hi, "bf16"), w, None, stride, pad)
    else:
        y = F.conv3d(x, w, None, stride, pad)
    y = y + b.view(1, -1, 1, 1, 1)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def trunk(sd, x, cfg):
    """Feature maps [stem, layer1..4] in fp32 *before* storage rounding plus the rounded tensors the next layer reads."""
    p = "r2plus1d_model."
    s, b = bn_affine(sd, p + "stem.1")
    h = F.conv3d(x, sd[p + "stem.0.weight"] * s.view(-1, 1, 1, 1, 1), b, (1, 2, 2), (0, 3, 3))
    h = rnd(F.relu(h), cfg.act)
    h = rnd(_conv(h, sd, p + "stem.3", p + "stem.4", 1, (1, 0, 0), cfg, name="stem.3"), cfg.act)
    feats = [h]
    skip = h                                              # the stem output is only ever stored in `act`
    for lname, _inp, _planes, stride in LAYERS:
        for blk in (0, 1):
            k = f"{p}{lname}.{blk}"
            st = stride if blk == 0 else 1
            t = rnd(_conv(h, sd, k + ".conv1.0.0", k + ".conv1.0.1", (1, st, st), (0, 1, 1), cfg, name=f"{lname}.{blk}.s1"), cfg.act)
            t = rnd(_conv(t, sd, k + ".conv1.0.3", k + ".conv1.1", (st, 1, 1), (1, 0, 0), cfg, name=f"{lname}.{blk}.t1"), cfg.act)
            t = rnd(_conv(t, sd, k + ".conv2.0.0", k + ".conv2.0.1", 1, (0, 1, 1), cfg, name=f"{lname}.{blk}.s2"), cfg.act)
            res = skip
            if blk == 0 and stride != 1:
                res = _conv(h, sd, k + ".downsample.0", k + ".downsample.1", (st, st, st), 0, cfg, relu=False, name=f"{lname}.{blk}.down")
                res = rnd(res, cfg.act if cfg.residual == "act" else "fp32")
            y = _conv(t, sd, k + ".conv2.0.3", k + ".conv2.1", 1, (1, 0, 0), cfg, residual=res, name=f"{lname}.{blk}.t2")
            h = rnd(y, cfg.act)
            skip = h if cfg.residual == "act" else y
        feats.append(h)
    return feats


def _axis(dst_size, in_size):
    """align_corners=True taps exactly as csrc axis_tap(): fp32 scale, truncation, l1 = src - i0, l0 = 1 - l1."""
    scale = torch.tensor((in_size - 1) / (dst_size - 1) if dst_size > 1 else 0.0, dtype=torch.float32)
    src = scale * torch.arange(dst_size, dtype=torch.float32)
    i0 = src.to(torch.int64).clamp_max(in_size - 1)
    i1 = i0 + (i0 < in_size - 1).to(torch.int64)
    l1 = src - i0.float()
    return i0, i1, 1.0 - l1, l1


def _interp_matrix(dst_size, in_size):
    i0, i1, l0, l1 = _axis(dst_size, in_size)
    m = torch.zeros(dst_size, in_size)
    m[torch.arange(dst_size), i0] += l0
    m[torch.arange(dst_size), i1] += l1 * (i1 != i0)      # a coincident second tap has weight 0 in the kernel too
    return m


def _hilo(v, kind):
    hi = rnd(v, kind)
    return hi, rnd(v - hi, kind)


def lateral_maps(sd, feats, t_out, cfg):
    """The four 64-channel maps the head reads: comb_1 (BN folded) applied per feature map at native resolution
    (stem + layer1 share a resolution and are summed), rounded to `cfg.lateral`, levels 2-4 interpolated along T."""
    s1, _t1 = bn_affine(sd, "comb_batch_norm_1")
    w1 = sd["comb_1_layer.weight"][:, :, 0, 0, 0] * s1.view(-1, 1)                  # (64, 1024)
    offs = [0]
    for f in feats:
        offs.append(offs[-1] + f.shape[1])
    g = []
    for i, f in enumerate(feats):
        w = w1[:, offs[i]:offs[i + 1]] if "lateral" in cfg.exact else rnd(w1[:, offs[i]:offs[i + 1]], cfg.wtrunk)
        g.append(torch.einsum("nctHW,oc->notHW", f, w))
    g = [g[0] + g[1]] + g[2:]
    g = [rnd(x, cfg.lateral) for x in g]
    for l in range(1, 4):
        i0, i1, l0, l1 = _axis(t_out, g[l].shape[2])
        a, b = g[l][:, :, i0], g[l][:, :, i1]
        up = l0.view(1, 1, -1, 1, 1) * a + l1.view(1, 1, -1, 1, 1) * b
        g[l] = rnd(torch.where((l1 == 0).view(1, 1, -1, 1, 1), a, up), cfg.lateral)
    return g


def head(sd, g, h_out, w_out, cfg, out_kind="logits"):
    """The fused head on lateral maps g (list of four (N,64,T,Hl,Wl) tensors at the output's frame rate)."""
    s1, t1 = bn_affine(sd, "comb_batch_norm_1")
    s2, t2 = bn_affine(sd, "comb_batch_norm_2")
    b1 = s1 * sd["comb_1_layer.bias"] + t1
    b2 = s2 * sd["comb_2_layer.bias"] + t2
    w2 = rnd(sd["comb_2_layer.weight"][:, :, 0, 0, 0] * s2.view(-1, 1), cfg.w2)
    wh = rnd(torch.cat([sd["segmentation_head.weight"], sd["motion_head.weight"]])[:, :, 0, 0, 0], cfg.wh)
    bh = torch.cat([sd["segmentation_head.bias"], sd["motion_head.bias"]])
    acc = 0.0
    wkind = "f16" if cfg.head == "row" else cfg.lateral
    for l in range(4):
        hl, wl = g[l].shape[3], g[l].shape[4]
        mh, mw = _interp_matrix(h_out, hl), _interp_matrix(w_out, wl)
        if cfg.head == "row":
            r = rnd(torch.einsum("hy,ncTyx->ncThx", mh, g[l]), "f16")                  # phase 1 -> fp16
            acc = acc + torch.einsum("wx,ncThx->ncThw", rnd(mw, "f16"), r)
        else:
            # patch head: A[(h,w),(y,x)] = rnd(wH[h,y] * wW[w,x]) - four non-zero entries per voxel and level
            ih = _axis(h_out, hl); iw = _axis(w_out, wl)
            for a in (0, 1):
                for b in (0, 1):
                    wgt = rnd(ih[2 + a].view(-1, 1) * iw[2 + b].view(1, -1), wkind)
                    if a == 1:
                        wgt = wgt * (ih[1] != ih[0]).view(-1, 1)
                    if b == 1:
                        wgt = wgt * (iw[1] != iw[0]).view(1, -1)
                    acc = acc + wgt.view(1, 1, 1, h_out, w_out) * g[l][:, :, :, ih[a]][:, :, :, :, iw[b]]
    bhi, blo = _hilo(b1, wkind)
    h1 = rnd(F.relu(acc + (bhi + blo).view(1, -1, 1, 1, 1)), cfg.h1)
    b2hi, b2lo = _hilo(b2, "bf16" if cfg.h1 == "bf16" else ("f16" if cfg.h1 == "f16" else "fp32"))
    h2 = rnd(F.relu(torch.einsum("ncThw,oc->noThw", h1, w2) + (b2hi + b2lo).view(1, -1, 1, 1, 1)), cfg.h2)
    o = torch.einsum("ncThw,oc->noThw", h2, wh) + bh.view(1, -1, 1, 1, 1)
    seg, mot = o[:, :2], torch.tanh(o[:, 2:])
    if out_kind == "prob":
        seg = torch.softmax(seg, 1)
    return rnd(seg, cfg.out), rnd(mot, cfg.out)


def decoder(sd, feats, t_out, h_out, w_out, cfg, out_kind="logits"):
    return head(sd, lateral_maps(sd, feats, t_out, cfg), h_out, w_out, cfg, out_kind)


def forward(sd, x, cfg=Config(), out_kind="logits"):
    """(seg logits or prob (N,2,T,H,W), tanh motion (N,4,T,H,W)) as fp32 tensors holding `cfg.out`-representable values."""
    sd = strip_module_prefix(sd)
    with torch.no_grad():
        feats = trunk(sd, x.float(), cfg)
        return decoder(sd, feats, x.shape[2], x.shape[3], x.shape[4], cfg, out_kind)
