"""clasfv_b200 - B200-native drop-in for CLAS-FV's full-video inference hot path.

Layout mirrors the reference so call sites keep reading the same:

    clasfv_b200.src.model.R2plus1D_18_MotionNet.R2plus1D_18_MotionNet   (reference: src/model/R2plus1D_18_MotionNet.py)
    clasfv_b200.src.fuse_utils.{divide_to_consecutive_clips, segment_a_video_with_fusion, compute_ef_using_putative_clips}
    clasfv_b200.src.transform_utils.generate_2dmotion_field
    clasfv_b200.src.echonet_dataset.{zeroone_normalizer, EDESpairs}
    clasfv_b200/motion_segment.py                                        (reference: motion_segment.py, same flags)

All device work goes through the C-ABI library ``csrc/libclasfv_b200.so`` (``include/clasfv_b200.h``),
hand-written sm_100a CUDA.  There is no CPU or PyTorch fallback: without the library, or
without a CUDA device, the operators raise.
"""
__version__ = "0.1.0"
PACKAGE_DIR = __import__("os").path.dirname(__import__("os").path.abspath(__file__))
