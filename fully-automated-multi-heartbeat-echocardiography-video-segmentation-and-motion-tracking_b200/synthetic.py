"""Synthetic inputs for tests and benchmarks: echo-like videos and random-init weights.

There is no network on the build or GPU boxes, so neither EchoNet-Dynamic videos nor the
authors' checkpoint are available; parity and throughput are measured on seeded synthetic
data of the reference's shapes (BASELINE.json ``configs``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

# (layer, inplanes, planes, stride) of torchvision's r2plus1d_18 trunk, as instantiated by the
# reference at src/model/R2plus1D_18_MotionNet.py:13
TRUNK_LAYERS = (("layer1", 64, 64, 1), ("layer2", 64, 128, 2), ("layer3", 128, 256, 2), ("layer4", 256, 512, 2))


def midplanes(inplanes: int, planes: int) -> int:
    return (inplanes * planes * 3 * 3 * 3) // (inplanes * 3 * 3 + 3 * planes)


def state_dict_spec():
    """[(key, shape, kind)] for the 242 tensors of the reference's state_dict, in its order.

    kind is one of conv / bn_weight / bn_bias / bn_mean / bn_var / bn_count / bias / fc.
    """
    spec = []

    def conv(key, cout, cin, k):
        spec.append((key + ".weight", (cout, cin) + k, "conv"))

    def bn(key, c):
        spec.append((key + ".weight", (c,), "bn_weight"))
        spec.append((key + ".bias", (c,), "bn_bias"))
        spec.append((key + ".running_mean", (c,), "bn_mean"))
        spec.append((key + ".running_var", (c,), "bn_var"))
        spec.append((key + ".num_batches_tracked", (), "bn_count"))

    p = "r2plus1d_model."
    conv(p + "stem.0", 45, 3, (1, 7, 7)); bn(p + "stem.1", 45)
    conv(p + "stem.3", 64, 45, (3, 1, 1)); bn(p + "stem.4", 64)
    for name, inplanes, planes, stride in TRUNK_LAYERS:
        for blk in (0, 1):
            cin = inplanes if blk == 0 else planes
            mid = midplanes(cin, planes)          # one value per block, shared by conv1 and conv2
            for cv, c_in in (("conv1", cin), ("conv2", planes)):
                key = f"{p}{name}.{blk}.{cv}"
                conv(key + ".0.0", mid, c_in, (1, 3, 3)); bn(key + ".0.1", mid)
                conv(key + ".0.3", planes, mid, (3, 1, 1)); bn(key + ".1", planes)
            if blk == 0 and stride != 1:
                conv(f"{p}{name}.0.downsample.0", planes, inplanes, (1, 1, 1)); bn(f"{p}{name}.0.downsample.1", planes)
    spec.append((p + "fc.weight", (400, 512), "fc"))
    spec.append((p + "fc.bias", (400,), "bias"))
    conv("comb_1_layer", 64, 1024, (1, 1, 1)); spec.append(("comb_1_layer.bias", (64,), "bias"))
    bn("comb_batch_norm_1", 64)
    conv("comb_2_layer", 64, 64, (1, 1, 1)); spec.append(("comb_2_layer.bias", (64,), "bias"))
    bn("comb_batch_norm_2", 64)
    conv("motion_head", 4, 64, (1, 1, 1)); spec.append(("motion_head.bias", (4,), "bias"))
    conv("segmentation_head", 2, 64, (1, 1, 1)); spec.append(("segmentation_head.bias", (2,), "bias"))
    return spec


def random_state_dict(seed: int = 0, prefix: str = ""):
    """Seeded random-init weights with the reference's 242 state_dict keys and shapes.

    Convolutions are He-normal (fan_out), BatchNorm affine / running statistics are drawn
    around the identity so that no layer is degenerate; the motion head keeps the reference's
    N(0, 1e-5) init (R2plus1D_18_MotionNet.py:23).  Deterministic for a given torch build.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in state_dict_spec():
        if kind == "conv":
            fan_out = shape[0] * int(np.prod(shape[2:]))
            std = math.sqrt(1e-5) if key.startswith("motion_head") else math.sqrt(2.0 / fan_out)
            t = torch.randn(shape, generator=g) * std
        elif kind == "bn_weight":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif kind == "bn_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind in ("bn_bias", "bn_mean"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_count":
            t = torch.tensor(1, dtype=torch.int64)
        elif kind == "fc":
            t = 0.01 * torch.randn(shape, generator=g)
        else:  # bias
            t = 0.05 * torch.randn(shape, generator=g)
        sd[prefix + key] = t
    return sd


def synthetic_echo_video(num_frames=128, height=112, width=112, seed=0, beats_per_clip=1.3, channels=3):
    """Seeded beating-ellipse + speckle video, float32 (channels, T, H, W) in [0, 1].

    A bright myocardium ring around a dark, periodically contracting cavity on a speckled
    sector background - enough structure for the network's activations to be non-trivial.
    All three channels are equal (grayscale echo, as after cv2 decode of a gray .avi).
    """
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    cy, cx = 0.55 * height, 0.5 * width
    frames = np.empty((num_frames, height, width), np.float32)
    speckle = rng.gamma(2.0, 0.08, size=(height, width)).astype(np.float32)
    for t in range(num_frames):
        phase = 2 * np.pi * beats_per_clip * t / 32.0
        a = 0.23 * height * (1.0 - 0.22 * (0.5 - 0.5 * np.cos(phase)))
        b = 0.15 * width * (1.0 - 0.30 * (0.5 - 0.5 * np.cos(phase)))
        r = np.sqrt(((yy - cy) / a) ** 2 + ((xx - cx) / b) ** 2)
        cavity = 1.0 / (1.0 + np.exp((r - 1.0) * 12.0))
        wall = np.exp(-((r - 1.25) ** 2) / 0.03)
        img = 0.15 + 0.55 * wall - 0.12 * cavity + speckle * (0.6 + 0.4 * rng.random((height, width), dtype=np.float32))
        frames[t] = img
    frames -= frames.min()
    frames /= frames.max()
    return np.repeat(frames[None], channels, axis=0).astype(np.float32)


def write_avi(path, video):
    """Write a (3, T, H, W) float [0,1] video as an .avi (for the motion_segment.py CLI tests)."""
    import cv2
    _, t, h, w = video.shape
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (w, h))
    if not wr.isOpened():
        raise RuntimeError("cv2.VideoWriter could not open " + str(path))
    for i in range(t):
        frame = (video[:, i].transpose(1, 2, 0)[..., ::-1] * 255.0).round().astype(np.uint8)
        wr.write(np.ascontiguousarray(frame))
    wr.release()
