"""Generate tests/golden/*.npz by running the UNMODIFIED reference code in the build container.

TEST INFRASTRUCTURE ONLY.  Usage (build container only; needs /root/reference):

    python oracle/make_golden.py

The reference has no tests and no golden vectors of its own, so the
oracle is pinned on outputs of the reference itself: each file stores the seeded inputs (or
the seeds that regenerate them) and what the reference's own function returned.  The vectors
are small (< 1 MB together) and committed; this script is the committed generator.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import clasfv_b200.synthetic as synthetic            # noqa: E402
from oracle import fixtures, ref_import              # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def stub_model(x):
    """Deterministic stand-in network for pinning the fusion control flow (no weights)."""
    lv = x[:, 0:1].float()
    lv = torch.nn.functional.avg_pool3d(lv, (1, 5, 5), 1, (0, 2, 2))
    seg = torch.cat([0.2 - lv, lv - 0.2], 1) * 12.0
    return seg, torch.zeros(x.shape[0], 4, *x.shape[2:])


def write_ef_golden(ref):
    from oracle import ef_ref
    out = {}
    for tag, (frames, period, seed) in {"a": (150, 47.0, 0), "b": (200, 61.0, 1), "c": (90, 33.0, 2)}.items():
        masks = ef_ref.beating_masks(frames, 112, period, seed)
        efs, pairs = ref.fuse_utils.compute_ef_using_putative_clips(masks, test_pat_index=tag, return_edes=True)
        out[f"args_{tag}"] = np.array([frames, period, seed], dtype=np.float64)
        out[f"efs_{tag}"] = np.array(efs, dtype=np.float64)
        out[f"pairs_{tag}"] = np.array(pairs, dtype=np.int64).reshape(-1, 2)
        length, radii = ref.echo_utils.get2dPucks((masks[5] == 1).astype("int"), (1.0, 1.0))
        out[f"pucks_{tag}"] = np.concatenate([[length], radii])
    np.savez_compressed(os.path.join(OUT, "ef.npz"), **out)


def main():
    ref = ref_import.import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)

    # 1. network forward, reference class, calibrated seeded weights, two small shapes
    sd = fixtures.calibrated_state_dict(0)
    net = ref.R2plus1D_18_MotionNet(pretrained=False)
    net.load_state_dict(sd)
    net.eval()
    out = {"weights_l1norm": np.float64(sum(float(v.double().abs().sum()) for v in sd.values()))}
    for tag, shape, seed in (("a", (8, 32, 32), 11), ("b", (16, 48, 32), 12)):
        x = fixtures.synthetic_clip(*shape, seed=seed)
        with torch.no_grad():
            seg, mot = net(x)
        out[f"x_{tag}"] = x.numpy()
        out[f"seg_{tag}"] = seg.numpy()
        out[f"motion_{tag}"] = mot.numpy()
    # one full-size clip, stored sub-sampled (every 8th pixel, every 4th frame)
    x = fixtures.synthetic_clip(32, 112, 112, seed=13)
    with torch.no_grad():
        seg, mot = net(x)
    out["seg_full_sub"] = seg.numpy()[:, :, ::4, ::8, ::8]
    out["motion_full_sub"] = mot.numpy()[:, :, ::4, ::8, ::8]
    out["seg_full_sum"] = np.float64(seg.double().sum())
    out["param_count"] = np.int64(sum(p.numel() for p in net.parameters() if p.requires_grad))
    np.savez_compressed(os.path.join(OUT, "model_forward.npz"), **out)

    # 2. warp primitive: generate_2dmotion_field + grid_sample (transform_utils.py:14-34)
    g = torch.Generator().manual_seed(3)
    src = torch.rand(2, 2, 12, 20, generator=g)
    flow = torch.tanh(0.15 * torch.randn(2, 2, 12, 20, generator=g))
    flow[0, :, 0, 0] = torch.tensor([-0.9, 0.9])        # exercise the border clamp
    with ref_import.cuda_is_noop():
        grid = ref.transform_utils.generate_2dmotion_field(src, flow)
        zero_grid = ref.transform_utils.generate_2dmotion_field(src, torch.zeros_like(flow))
    warped = torch.nn.functional.grid_sample(src, grid, align_corners=False, mode="bilinear", padding_mode="border")
    colimg = torch.arange(112.0).view(1, 1, 1, 112).expand(1, 1, 112, 112).contiguous()
    with ref_import.cuda_is_noop():
        zg = ref.transform_utils.generate_2dmotion_field(colimg, torch.zeros(1, 2, 112, 112))
    zero_flow_cols = torch.nn.functional.grid_sample(colimg, zg, align_corners=False, padding_mode="border")[0, 0, 0]
    np.savez_compressed(os.path.join(OUT, "warp.npz"), src=src.numpy(), flow=flow.numpy(), grid=grid.numpy(),
                        zero_grid=zero_grid.numpy(), warped=warped.numpy(), zero_flow_cols=zero_flow_cols.numpy())

    # 3. divide_to_consecutive_clips (fuse_utils.py:16-33) - hard-codes 112x112
    out = {}
    for length in (75, 48, 64, 80):
        video = synthetic.synthetic_echo_video(length, 112, 112, seed=20 + length)
        clips = ref.fuse_utils.divide_to_consecutive_clips(video, interpolate_last=True)
        out[f"shape_{length}"] = np.array(clips.shape)
        out[f"sub_{length}"] = clips[:, :, :, ::16, ::16]
        out[f"sum_{length}"] = np.float64(clips.sum())
        out[f"dtype_{length}"] = np.array(str(clips.dtype))
    np.savez_compressed(os.path.join(OUT, "divide_clips.npz"), **out)

    # 4. segment_a_video_with_fusion control flow (fuse_utils.py:36-102), stub network + majority voter.
    #    T=80: every shift has round(L/32)=2 clips, so the reference's np.array(list) is not ragged
    #    and the function runs unmodified on NumPy 2.x.
    out = {}
    for tag, length, f, step in (("f5", 80, 5, 1), ("f1", 64, 1, 1), ("f12", 80, 12, 1)):
        video = synthetic.synthetic_echo_video(length, 112, 112, seed=40 + f)
        fused = ref.fuse_utils.segment_a_video_with_fusion(video, stub_model, interpolate_last=True, step=step,
                                                           num_clips=f, fuse_method="simple", class_list=[0, 1])
        out[f"shape_{tag}"] = np.array(fused.shape)
        out[f"dtype_{tag}"] = np.array(str(fused.dtype))
        out[f"bits_{tag}"] = np.packbits(fused.astype(np.uint8))
        out[f"args_{tag}"] = np.array([length, f, step, 40 + f])
    np.savez_compressed(os.path.join(OUT, "fusion_flow.npz"), **out)

    # 5. small host helpers
    rng = np.random.default_rng(5)
    v = (rng.random((3, 4, 6, 5)) * 200 + 13).astype(np.float32)
    norm = ref.echonet_dataset.zeroone_normalizer(v.copy())
    pairs = ref.echonet_dataset.EDESpairs([0, 31, 62, 95], [14, 47, 49, 80, 120])
    np.savez_compressed(os.path.join(OUT, "host_helpers.npz"), v=v, norm=norm, pairs=np.array(pairs))

    # 6. ejection fraction: the reference's compute_ef_using_putative_clips + get2dPucks (find_boundaries bound to the oracle's
    # restatement, the only substitution) on seeded synthetic multi-heartbeat masks
    write_ef_golden(ref)
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
