"""Drop-in for the reference's ``src/model/R2plus1D_18_MotionNet.py`` running on libclasfv_b200.

Same constructor, same 242-key ``state_dict`` (so ``load_state_dict(torch.load(path)["model"])`` and
``torch.nn.DataParallel(...)`` from ``motion_segment.py:69-76`` keep working), same
``forward(x) -> (segmentation_output, motion_output)`` (reference ``:26-71``).  The ``nn.Module`` tree
(torchvision's ``r2plus1d_18`` trunk + the five decoder layers, built exactly as the reference's
``__init__`` ``:11-24`` does) only *holds* the parameters; every FLOP of ``forward`` runs in the
C-ABI library's sm_100a kernels.  There is no PyTorch or CPU execution path: inference mode on a
CUDA device or an error.

Extra, optional: ``precision`` - constructor keyword or attribute: "fp32" reference-tolerance mode on CUDA cores;
"bf16" / "fp16" tcgen05 tensor-core modes (same kernels, same speed; fp16 stores 11 significant bits instead of 8 and
is the 16-bit mode that meets the parity gates).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from ... import engine as _engine
from ..._lib import OUT_LOGITS, OUT_PROB, ClasfvError


class R2plus1D_18_MotionNet(nn.Module):
    def __init__(self, pretrained=True, output_channels=4, precision="fp32"):
        super(R2plus1D_18_MotionNet, self).__init__()
        if output_channels != 4:
            raise ClasfvError("the motion head has 4 channels [fwd x, fwd y, bwd x, bwd y] (reference :22, :67)")
        from torchvision.models.video import r2plus1d_18
        if pretrained:
            # reference :13 downloads the Kinetics-400 trunk; same behaviour (raises when offline)
            from torchvision.models.video import R2Plus1D_18_Weights
            self.r2plus1d_model = r2plus1d_18(weights=R2Plus1D_18_Weights.KINETICS400_V1)
        else:
            self.r2plus1d_model = r2plus1d_18(weights=None)
        self.comb_1_layer = nn.Conv3d(1024, 64, 1)
        self.comb_batch_norm_1 = nn.BatchNorm3d(64)
        self.comb_relu_1 = nn.ReLU(inplace=True)
        self.comb_2_layer = nn.Conv3d(64, 64, 1)
        self.comb_batch_norm_2 = nn.BatchNorm3d(64)
        self.comb_relu_2 = nn.ReLU(inplace=True)
        self.motion_head = nn.Conv3d(64, 4, 1)
        nn.init.normal_(self.motion_head.weight, mean=0.0, std=np.sqrt(1e-5))
        self.segmentation_head = nn.Conv3d(64, 2, 1)
        self.precision = precision
        self._engines = {}          # device index -> (Engine, signature)
        self._packed_from = None    # set on DataParallel replicas: the module whose parameters the cache key follows

    # ------------------------------------------------------------------ engine management
    def _signature(self):
        # a DataParallel replica receives freshly broadcast parameter tensors on every forward: its cache key is the SOURCE
        # module's (whose tensors persist), otherwise every call would repack and re-upload the whole network
        src = getattr(self, "_packed_from", None) or self
        sig = [self.precision]
        for t in list(src.parameters()) + list(src.buffers()):
            sig.append((t.data_ptr(), t._version))
        return tuple(sig)

    def engine(self, device=None):
        """The packed network for ``device`` (default: where the parameters live), repacked whenever
        a parameter / buffer changed (load_state_dict, .to(), in-place updates) or precision changed."""
        pdev = next(self.parameters()).device
        dev = torch.device(device) if device is not None else pdev
        if dev.type != "cuda":
            raise ClasfvError("R2plus1D_18_MotionNet (clasfv_b200) runs on CUDA only - move the model with "
                              ".to('cuda'); there is no CPU path")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        sig = self._signature()
        cached = self._engines.get(idx)
        if cached is None or cached[1] != sig:
            eng = cached[0] if cached is not None else _engine.Engine(torch.device("cuda", idx))
            eng.load_state_dict(self.state_dict(), self.precision)
            self._engines[idx] = (eng, sig)
        return self._engines[idx][0]

    def __getstate__(self):             # engines hold C handles: never pickle / replicate them
        state = self.__dict__.copy()
        state["_engines"] = {}
        state["_packed_from"] = None
        return state

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica._engines = self._engines
        replica._packed_from = getattr(self, "_packed_from", None) or self
        return replica

    # ------------------------------------------------------------------ forward
    def _prepare(self, x):
        if self.training:
            raise ClasfvError("clasfv_b200 implements inference only (BatchNorm running statistics): call model.eval()")
        if not isinstance(x, torch.Tensor) or x.dim() != 5 or x.shape[1] != 3:
            raise ClasfvError("expected a float tensor of shape (N,3,T,H,W)")
        pdev = next(self.parameters()).device
        if pdev.type != "cuda":
            raise ClasfvError("model parameters are on the CPU; clasfv_b200 has no CPU path (use .to('cuda'))")
        return x.to(device=pdev, dtype=torch.float32, non_blocking=True).contiguous()

    @torch.no_grad()
    def forward(self, x):
        x = self._prepare(x)
        return self.engine(x.device).forward(x, OUT_LOGITS, torch.float32)

    @torch.no_grad()
    def forward_prob(self, x, out_dtype=None):
        """forward + the F.softmax(seg, 1) of fuse_utils.py:60 fused into the head kernel."""
        x = self._prepare(x)
        if out_dtype is None:
            out_dtype = _engine.storage_dtype(self.precision)
        return self.engine(x.device).forward(x, OUT_PROB, out_dtype)
