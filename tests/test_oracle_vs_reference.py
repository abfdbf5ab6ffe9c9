"""Live pin of the oracle against the reference checkout (build container only; skipped on
the GPU box, where /root/reference does not exist)."""
import numpy as np
import pytest
import torch

import clasfv_b200.synthetic as synthetic
from oracle import fixtures, fuse_ref, model_ref, ref_import
from oracle.make_golden import stub_model

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return ref_import.import_reference()


def test_state_dict_spec_matches_reference(ref):
    sd = ref.R2plus1D_18_MotionNet(pretrained=False).state_dict()
    spec = synthetic.state_dict_spec()
    assert [k for k, _s, _k in spec] == list(sd.keys())
    for k, shape, _kind in spec:
        assert tuple(sd[k].shape) == tuple(shape)


def test_forward_bit_identical_to_reference_class(ref):
    sd = synthetic.random_state_dict(3)
    net = ref.R2plus1D_18_MotionNet(pretrained=False)
    net.load_state_dict(sd)
    net.eval()
    x = torch.rand(2, 3, 16, 32, 48, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        seg, mot = net(x)
    seg2, mot2 = model_ref.forward(sd, x)
    assert torch.equal(seg, seg2) and torch.equal(mot, mot2)


def test_fusion_matches_reference_function(ref):
    video = synthetic.synthetic_echo_video(80, 112, 112, seed=77)
    a = ref.fuse_utils.segment_a_video_with_fusion(video, stub_model, interpolate_last=True, step=1, num_clips=7)
    b = fuse_ref.segment_a_video_with_fusion(video, stub_model, interpolate_last=True, step=1, num_clips=7)
    assert a.dtype == b.dtype and np.array_equal(a, b)


def test_warp_matches_reference_function(ref):
    g = torch.Generator().manual_seed(9)
    src = torch.rand(1, 2, 20, 28, generator=g)
    flow = torch.tanh(0.2 * torch.randn(1, 2, 20, 28, generator=g))
    with ref_import.cuda_is_noop():
        grid = ref.transform_utils.generate_2dmotion_field(src, flow)
    assert torch.equal(grid, fuse_ref.generate_2dmotion_field(src, flow))


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
@pytest.mark.parametrize("forward", [True, False])
def test_apply_sequence_deformation_matches_reference_function(ref, mode, forward):
    """The motion-tracking product (src/visualization_utils.py:106-128), unmodified reference function with ``.cuda()``
    patched to a no-op, against the oracle's restatement: bit-identical."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src import visualization_utils as ref_vis
    g = torch.Generator().manual_seed(4)
    src = torch.rand(2, 1, 24, 32, generator=g)
    motion = torch.tanh(0.1 * torch.randn(2, 4, 10, 24, 32, generator=g))
    a, b = (1, 6) if forward else (8, 3)
    with ref_import.cuda_is_noop():
        want = ref_vis.apply_sequence_deformation(src, motion, a, b, grid_mode=mode, forward=forward)
    got = fuse_ref.apply_sequence_deformation(src, motion, a, b, grid_mode=mode, forward=forward)
    assert torch.equal(want, got)


def test_ef_matches_reference_functions(ref):
    """oracle/ef_ref.py and the product's host EF code against the reference's compute_ef_using_putative_clips, EDESpairs and
    get2dPucks run unmodified (find_boundaries bound to the restatement on both sides - skimage is absent)."""
    from clasfv_b200.src import fuse_utils
    from oracle import ef_ref
    for frames, period, seed in ((130, 41.0, 7), (180, 52.0, 8)):
        masks = ef_ref.beating_masks(frames, 112, period, seed)
        want, want_pairs = ref.fuse_utils.compute_ef_using_putative_clips(masks, test_pat_index="t", return_edes=True)
        assert len(want) >= 2
        for impl in (ef_ref.compute_ef_using_putative_clips, fuse_utils.compute_ef_using_putative_clips):
            got, pairs = impl(masks, "t", return_edes=True)
            assert [tuple(map(int, p)) for p in pairs] == [tuple(map(int, p)) for p in want_pairs]
            np.testing.assert_allclose(np.array(got), np.array(want), rtol=1e-12)
        for f in (0, 11, 60):
            lw, rw = ref.echo_utils.get2dPucks((masks[f] == 1).astype("int"), (1.0, 1.0))
            for pucks in (ef_ref.get_2d_pucks, fuse_utils.get2dPucks):
                lg, rg = pucks((masks[f] == 1).astype("int"), (1.0, 1.0))
                np.testing.assert_allclose(np.concatenate([[lg], rg]), np.concatenate([[lw], rw]), rtol=1e-12)
    d, s_ = [5, 40, 90, 91], [20, 22, 60, 130, 3]
    assert fuse_utils.EDESpairs(d, s_) == ref.echonet_dataset.EDESpairs(d, s_) == ef_ref.edes_pairs(d, s_)
