"""Import the UNMODIFIED reference modules from /root/reference in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` and by the live pin tests
(skipped when the checkout is absent, as on the GPU box).  The reference modules import
packages that are not installed here and cannot be (no network): SimpleITK, LabelFusion,
echonet, skimage, matplotlib, IPython, h5py.  None of them is *called* on the code paths
we pin except the voter and the SimpleITK array wrappers, so they are replaced in
``sys.modules`` by inert stubs; ``fuse_images`` is bound to the oracle's majority voter
and the two SimpleITK array converters to identity copies.  ``Tensor.cuda`` is patched to
a no-op so ``generate_2dmotion_field`` (hard-coded ``.cuda()``, transform_utils.py:19-20)
runs on the CPU.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("CLASFV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "fuse_utils.py"))


class _Stub(types.ModuleType):
    """A module whose every attribute is an inert class (so ``from x import y`` works)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


_MISSING = ("SimpleITK", "LabelFusion", "LabelFusion.wrapper", "echonet", "echonet.datasets",
            "skimage", "skimage.transform", "skimage.segmentation", "skimage.measure", "skimage.morphology",
            "skimage.filters", "skimage.exposure", "skimage.util",
            "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors",
            "matplotlib.figure", "matplotlib.backends", "matplotlib.backends.backend_agg",
            "IPython", "IPython.display", "h5py")


def install_stubs(voter=None):
    from oracle import fuse_ref
    for name in _MISSING:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    itk = sys.modules["SimpleITK"]
    if isinstance(itk, _Stub):
        itk.GetImageFromArray = lambda arr, isVector=False: np.array(arr, copy=True)
        itk.GetArrayFromImage = lambda img: np.asarray(img)
    lf = sys.modules["LabelFusion.wrapper"]
    if isinstance(lf, _Stub):
        vote = voter or fuse_ref.majority_vote
        lf.fuse_images = lambda images, method="simple", class_list=None: vote(images)
    seg = sys.modules["skimage.segmentation"]
    if isinstance(seg, _Stub):
        # the one skimage function on the EF path (get2dPucks, src/utils/echo_utils.py:301): bound to the oracle's restatement
        from oracle import ef_ref
        seg.find_boundaries = lambda label_img, mode="thick", **kw: ef_ref.find_boundaries_thick(label_img)


def import_reference():
    """Returns a namespace with the reference's hot-path callables."""
    if not available():
        raise FileNotFoundError(REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
        from src import fuse_utils, transform_utils, echonet_dataset, clasfv_losses
        from src.utils import echo_utils
    return types.SimpleNamespace(
        R2plus1D_18_MotionNet=R2plus1D_18_MotionNet, fuse_utils=fuse_utils, transform_utils=transform_utils,
        echonet_dataset=echonet_dataset, clasfv_losses=clasfv_losses, echo_utils=echo_utils)


@contextlib.contextmanager
def cuda_is_noop():
    import torch
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = saved
