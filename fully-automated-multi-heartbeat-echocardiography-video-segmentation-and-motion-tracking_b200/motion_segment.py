"""Drop-in for the reference's ``motion_segment.py`` CLI (same flags, same output files) on clasfv_b200.

    python motion_segment.py -p video.avi -m model.pth -d cuda -f 32 -s 1 -c all -o out/

Reference flags kept (motion_segment.py:19-65): -p/--path, -m/--model, -d/--device, --fuse_method,
-f/--fuse, -s/--step, -o/--output, -v/--verbose, -c/--content, --height, --width.
Added: --precision {fp32,bf16,fp16}; --fuse_method also accepts "warp" (the warp-and-fuse operator).
``-d cpu`` (the reference's default) is rejected: this build has no CPU path.
"""
import argparse
import os
import pickle
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import clasfv_b200  # noqa: E402,F401
from clasfv_b200.src.fuse_utils import compute_ef_using_putative_clips, segment_a_video_with_fusion  # noqa: E402
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet  # noqa: E402


def build_parser():
    ap = argparse.ArgumentParser(description="Segment and motion track heart structure in an Echo Video")
    ap.add_argument("-p", "--path", required=True, type=str, help="Path to the video")
    ap.add_argument("-m", "--model", required=False, type=str, help="Path to the saved model weights",
                    default="save_models/R2plus1DMotionSegNet_model.pth")
    ap.add_argument("-d", "--device", required=False, type=str, help="Which device to use: CPU or GPU", default="cpu")
    ap.add_argument("--fuse_method", required=False, type=str, help="Fuse method", default="simple")
    ap.add_argument("-f", "--fuse", required=False, type=int, help="Number of shifted video clips to fuse", default=1)
    ap.add_argument("-s", "--step", required=False, type=int, help="Step of shifting", default=1)
    ap.add_argument("-o", "--output", required=False, type=str, help="Path to the output files", default=".")
    ap.add_argument("-v", "--verbose", action='store_true', help="Verbosity")
    ap.add_argument("-c", "--content", required=False, type=str,
                    help="Content of the output: gif, binary, binary_video, all", default="binary")
    ap.add_argument("--height", required=False, type=int, help="Height of image (pretrain model uses 112)", default=112)
    ap.add_argument("--width", required=False, type=int, help="Width of image (pretrain model uses 112)", default=112)
    ap.add_argument("--precision", required=False, type=str, choices=("fp32", "bf16", "fp16"), default="fp32",
                    help="fp32: reference-tolerance mode; bf16 / fp16: tensor-core modes")
    return ap


def load_frames(path):
    """cv2 decode -> (T, H, W, 3) uint8 frames in cv2's B,G,R byte order (motion_segment.py:80-92 without the per-frame
    cvtColor: the channel swap happens in the device ingest)."""
    import cv2
    capture = cv2.VideoCapture(path)
    frame_count = int(capture.get(cv2.CAP_PROP_FRAME_COUNT))
    frame_width = int(capture.get(cv2.CAP_PROP_FRAME_WIDTH))
    frame_height = int(capture.get(cv2.CAP_PROP_FRAME_HEIGHT))
    video = np.zeros((frame_count, frame_height, frame_width, 3), np.uint8)
    for count in range(frame_count):
        ret, frame = capture.read()
        if not ret:
            raise ValueError("Failed to load frame #{} of {}.".format(count, path))
        video[count, :, :] = frame
    return video


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.device.lower().startswith("cpu") or not torch.cuda.is_available():
        sys.exit("clasfv_b200: this build runs on CUDA (sm_100a) only - pass `-d cuda`. There is no CPU path; "
                 "the reference's PyTorch CPU path is the oracle/baseline, not part of this package.")
    model = torch.nn.DataParallel(R2plus1D_18_MotionNet(pretrained=False, precision=args.precision), device_ids=[torch.device(args.device).index or 0])
    model.to(args.device)
    model.load_state_dict(torch.load(args.model, map_location=args.device)["model"])
    if args.verbose:
        print(f'R2+1D MotionNet has {sum(p.numel() for p in model.parameters() if p.requires_grad)} parameters.')
    model.eval()

    # decode on the host, everything after it on the device: channel swap, trilinear pre-resize (align_corners=True)
    # and zero-one normalisation are one ingest call (clasfv_ingest_u8); the (3,T,h,w) video never exists on the host
    eng = (model.module if isinstance(model, torch.nn.DataParallel) else model).engine()
    video = eng.ingest_u8(load_frames(args.path), args.height, args.width, bgr=True)
    class_list = [0, 1]
    segmentations = segment_a_video_with_fusion(video, model=model, interpolate_last=True, step=args.step,
                                                num_clips=args.fuse, fuse_method=args.fuse_method, class_list=class_list)
    predicted_efs, edes_pairs = compute_ef_using_putative_clips(segmentations, test_pat_index=args.path, return_edes=True)

    if args.verbose:
        print("Identified {:d} systoles".format(len(predicted_efs)))
        if len(predicted_efs) > 0:
            print("\nEjection fractions measured at each systole are:")
            for i in range(len(predicted_efs)):
                print("Systole #{:d}: ED {:d} & ES {:d} length={:d}".format(i + 1, edes_pairs[i][0], edes_pairs[i][1],
                                                                            edes_pairs[i][1] - edes_pairs[i][0]))
                print("EF: {:.2f}\n".format(predicted_efs[i]))
            print("The average ejection fraction is {:.2f}".format(np.mean(predicted_efs)))

    filename = args.path[args.path.rfind("/") + 1:args.path.rfind(".")]
    content = args.content.lower().split(",")
    os.makedirs(args.output, exist_ok=True)
    if "gif" in content or "all" in content:
        try:
            from clasfv_b200.src.visualization_utils import make_annotated_gif
            make_annotated_gif(segmentations, video.cpu().numpy(), filename=os.path.join(args.output, filename + "_annotated.gif"))
        except ImportError as e:      # matplotlib is optional presentation tooling
            print("skipping the annotated gif:", e)
    if "binary" in content or "all" in content:
        for i in range(len(edes_pairs)):
            ed_index, es_index = edes_pairs[i][0], edes_pairs[i][1]
            with open(os.path.join(args.output, filename + "_ED_Frame_{:d}_segmentation.pkl".format(ed_index)), "wb") as outfile:
                pickle.dump(segmentations[ed_index], outfile)
            with open(os.path.join(args.output, filename + "_ES_Frame_{:d}_segmentation.pkl".format(es_index)), "wb") as outfile:
                pickle.dump(segmentations[es_index], outfile)
    if "binary_video" in content or "all" in content:
        with open(os.path.join(args.output, filename + "_whole_video_segmentation.pkl"), "wb") as outfile:
            pickle.dump(segmentations, outfile)
    return 0


if __name__ == "__main__":
    sys.exit(main())
