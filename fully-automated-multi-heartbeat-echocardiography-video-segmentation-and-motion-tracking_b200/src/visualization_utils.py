"""Drop-ins for two functions of the reference's ``src/visualization_utils.py``:

* ``apply_sequence_deformation`` (:106-128) - the motion-tracking product: an image or label of one frame is carried
  through a run of the clip's forward / backward motion fields by chained warps (label propagation from ED to ES and
  back).  Same signature; every warp is the library's ``clasfv_warp_mode`` kernel (no PyTorch arithmetic).
* ``make_annotated_gif`` (:476-538) - presentation only, off the hot path.  Written without matplotlib: the LV mask is
  blended over the grayscale frame and frames are written as an animated GIF via Pillow when available."""
import numpy as np

from .transform_utils import warp as _warp


def apply_sequence_deformation(flow_source_image, motion_output, start_index, end_index, grid_mode="bilinear", forward=True):
    """flow_source_image (N,C,H,W) CUDA: image or label of frame ``start_index``; motion_output (N,4,T,H,W) from the
    motion head ([fwd x, fwd y, bwd x, bwd y]).  Applies the deformations of frames ``range(start_index, end_index, +-1)``
    (forward fields when ``forward`` else backward fields) one after the other; ``grid_mode`` "nearest" for labels,
    "bilinear" for images, as the reference recommends.  Like the reference it raises UnboundLocalError for an empty range."""
    step = 1 if forward else -1
    new_image = None
    for frame_index in range(start_index, end_index, step):
        field = motion_output[:, :2, frame_index] if forward else motion_output[:, 2:, frame_index]
        new_image = _warp(flow_source_image, field, grid_mode)
        flow_source_image = new_image
    if new_image is None:
        raise UnboundLocalError("local variable 'new_image' referenced before assignment (empty frame range)")
    return new_image


def make_annotated_gif(segmentations, video, filename="annotated.gif", alpha=0.4, fps=30):
    try:
        from PIL import Image
    except ImportError as e:
        raise ImportError("writing a gif needs Pillow") from e
    frames = []
    vid = np.asarray(video)
    for t in range(min(len(segmentations), vid.shape[1])):
        rgb = np.clip(vid[:, t].transpose(1, 2, 0), 0, 1).copy()
        m = np.asarray(segmentations[t]) == 1
        rgb[m] = (1 - alpha) * rgb[m] + alpha * np.array([0.1, 0.4, 1.0])
        frames.append(Image.fromarray((rgb * 255).astype(np.uint8)))
    frames[0].save(filename, save_all=True, append_images=frames[1:], duration=int(1000 / fps), loop=0)
    return filename
