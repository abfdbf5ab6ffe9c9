"""``make_annotated_gif`` of the reference's ``src/visualization_utils.py:476-538`` (presentation only,
off the hot path).  Written with cv2 alone (matplotlib is not required): the LV mask is blended over the
grayscale frame and frames are written as an animated GIF via Pillow when available."""
import numpy as np


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
