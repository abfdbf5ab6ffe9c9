"""Oracle: functional restatement of the CLAS-FV network forward pass.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, on top of plain ``torch.nn.functional`` CPU ops and a reference-keyed
``state_dict`` (keys without the ``module.`` DataParallel prefix):

* ``src/model/R2plus1D_18_MotionNet.py:26-71``  (forward: trunk, 5 trilinear
  upsamples with align_corners=True, channel concat, comb_1/comb_2 1x1x1
  conv + BN + ReLU, segmentation head, motion head + tanh)
* the third-party trunk the reference instantiates at
  ``src/model/R2plus1D_18_MotionNet.py:13`` -
  ``torchvision.models.video.r2plus1d_18`` (pinned 0.6.0 in the reference's
  requirements.txt:161; 0.26.0 in this image): R2Plus1dStem, four layers of two
  BasicBlocks, each conv a (1x3x3 spatial -> BN -> ReLU -> 3x1x1 temporal)
  factorisation, 1x1x1 strided downsample + BN on the first block of layers
  2-4.  BN is evaluated in inference mode (running statistics, eps 1e-5).

``calibrate=True`` additionally *rewrites* every BatchNorm's running statistics
with the statistics of the activations it sees (used by the test fixtures to
produce a well-conditioned random network).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5

# (name, inplanes, planes, stride) for layer1..layer4 of r2plus1d_18
LAYERS = (("layer1", 64, 64, 1), ("layer2", 64, 128, 2),
          ("layer3", 128, 256, 2), ("layer4", 256, 512, 2))
# decoder up-sampling factors, src/model/R2plus1D_18_MotionNet.py:41-49
UPSAMPLE = ((1, 2, 2), (1, 2, 2), (2, 4, 4), (4, 8, 8), (8, 16, 16))


def midplanes(inplanes: int, planes: int) -> int:
    """Channel count of the (2+1)D factorisation's intermediate tensor."""
    return (inplanes * planes * 3 * 3 * 3) // (inplanes * 3 * 3 + 3 * planes)


def strip_module_prefix(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def _bn(x, sd, key, calibrate):
    if calibrate:
        red = (0, 2, 3, 4)
        sd[key + ".running_mean"] = x.mean(red).detach().clone()
        sd[key + ".running_var"] = x.var(red, unbiased=False).detach().clone().clamp_min(1e-6)
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                        sd[key + ".weight"], sd[key + ".bias"], False, 0.0, BN_EPS)


def _conv2plus1d(x, sd, key, stride, calibrate):
    # spatial 1x3x3 (stride (1,s,s), pad (0,1,1)) -> BN -> ReLU -> temporal 3x1x1 (stride (s,1,1), pad (1,0,0))
    x = F.conv3d(x, sd[key + ".0.weight"], None, (1, stride, stride), (0, 1, 1))
    x = F.relu(_bn(x, sd, key + ".1", calibrate))
    return F.conv3d(x, sd[key + ".3.weight"], None, (stride, 1, 1), (1, 0, 0))


def _basic_block(x, sd, key, stride, has_down, calibrate):
    out = _conv2plus1d(x, sd, key + ".conv1.0", stride, calibrate)
    out = F.relu(_bn(out, sd, key + ".conv1.1", calibrate))
    out = _conv2plus1d(out, sd, key + ".conv2.0", 1, calibrate)
    out = _bn(out, sd, key + ".conv2.1", calibrate)
    res = x
    if has_down:
        res = F.conv3d(x, sd[key + ".downsample.0.weight"], None, (stride,) * 3, 0)
        res = _bn(res, sd, key + ".downsample.1", calibrate)
    return F.relu(out + res)


def trunk_features(sd, x, calibrate=False):
    """Stem + layer1..4 feature maps (R2plus1D_18_MotionNet.py:29-37)."""
    p = "r2plus1d_model."
    s = F.conv3d(x, sd[p + "stem.0.weight"], None, (1, 2, 2), (0, 3, 3))
    s = F.relu(_bn(s, sd, p + "stem.1", calibrate))
    s = F.conv3d(s, sd[p + "stem.3.weight"], None, 1, (1, 0, 0))
    s = F.relu(_bn(s, sd, p + "stem.4", calibrate))
    feats = [s]
    h = s
    for name, _inp, _planes, stride in LAYERS:
        h = _basic_block(h, sd, p + name + ".0", stride, stride != 1, calibrate)
        h = _basic_block(h, sd, p + name + ".1", 1, False, calibrate)
        feats.append(h)
    return feats


def forward(sd, x, calibrate=False):
    """(segmentation logits (N,2,T,H,W), tanh motion (N,4,T,H,W)).

    Follows src/model/R2plus1D_18_MotionNet.py:26-71 line by line.
    """
    sd = sd if calibrate else strip_module_prefix(sd)
    with torch.no_grad():
        feats = trunk_features(sd, x, calibrate)
        ups = [F.interpolate(f, scale_factor=list(sf), mode="trilinear", align_corners=True)
               for f, sf in zip(feats, UPSAMPLE)]
        cat = torch.cat(ups, 1)                                              # :52
        h = F.conv3d(cat, sd["comb_1_layer.weight"], sd["comb_1_layer.bias"])  # :55
        h = F.relu(_bn(h, sd, "comb_batch_norm_1", calibrate))               # :56-57
        h = F.conv3d(h, sd["comb_2_layer.weight"], sd["comb_2_layer.bias"])  # :60
        h = F.relu(_bn(h, sd, "comb_batch_norm_2", calibrate))               # :61-62
        seg = F.conv3d(h, sd["segmentation_head.weight"], sd["segmentation_head.bias"])  # :65
        mot = torch.tanh(F.conv3d(h, sd["motion_head.weight"], sd["motion_head.bias"]))  # :68-69
    return seg, mot


def decoder_features(sd, x):
    """The 64-channel tensor feeding both heads (for calibrating head scales)."""
    sd = strip_module_prefix(sd)
    with torch.no_grad():
        feats = trunk_features(sd, x)
        ups = [F.interpolate(f, scale_factor=list(sf), mode="trilinear", align_corners=True)
               for f, sf in zip(feats, UPSAMPLE)]
        h = F.conv3d(torch.cat(ups, 1), sd["comb_1_layer.weight"], sd["comb_1_layer.bias"])
        h = F.relu(_bn(h, sd, "comb_batch_norm_1", False))
        h = F.conv3d(h, sd["comb_2_layer.weight"], sd["comb_2_layer.bias"])
        return F.relu(_bn(h, sd, "comb_batch_norm_2", False))
