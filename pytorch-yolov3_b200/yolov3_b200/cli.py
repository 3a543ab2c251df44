"""`yolov3` command line on top of the B200 hot path — the batched frame loops SURVEY.md §8f #4 asks
for: same flags and semantics as the reference's CLI (yolov3/__main__.py:36-96: ``-C|-I|-V``, ``-c``,
``-w``, ``-d``, ``-i 0.3``, ``-p 0.05``, ``-n``, ``-o``, ``--show-fps``, ``-v``), but the image-directory
mode and the video mode push BATCHES through ``inference_batches`` (the reference processes one image
per call — ``# TODO: batch images``, yolov3/__main__.py:157 — and one frame per call,
yolov3/inference.py:522-541), so upload, kernels and download of consecutive batches overlap.

Display / drawing / webcam threads are not part of the hot path: when the reference package is
importable (``baseline/_ref`` or an installed ``pytorch-yolov3``) its ``draw_boxes`` / ``detect_in_cam``
are used with the hot path re-bound to this package (INTEGRATION.md recipe b); otherwise boxes are
drawn with plain ``cv2.rectangle`` and the webcam mode is unavailable.  ``--no-display`` and
``--save-json`` (not in the reference) make the CLI usable on a headless GPU box.
"""
import argparse
import importlib
import json
import os
import pathlib
import sys
import time

import numpy as np

from . import Darknet, inference, inference_batches, non_max_suppression


def reference_package():
    """The reference's `yolov3` package with its hot path re-bound to this implementation, or None.
    (INTEGRATION.md recipe b; looks at an installed package first, then <repo>/baseline/_ref.)"""
    here = os.path.dirname(os.path.abspath(__file__))
    candidates = [None, os.path.join(os.path.dirname(os.path.dirname(here)), "baseline", "_ref")]
    for extra in candidates:
        saved_path, saved_mod = list(sys.path), sys.modules.pop("yolov3", None)
        try:
            if extra is not None:
                if not os.path.isdir(os.path.join(extra, "yolov3")):
                    continue
                sys.path.insert(0, extra)
            # the alias package that ships with this repo is also called `yolov3`: skip it
            sys.path[:] = [p for p in sys.path if not os.path.isfile(os.path.join(p or ".", "yolov3", "_b200_alias"))]
            np.int = int  # the reference uses the alias NumPy removed (yolov3/inference.py:353)
            ref = importlib.import_module("yolov3")
            if not hasattr(ref, "draw_boxes"):
                raise ImportError("not the reference package")
            ref_inf = importlib.import_module("yolov3.inference")
            ref.Darknet, ref.inference, ref.non_max_suppression = Darknet, inference, non_max_suppression
            ref_inf.inference, ref_inf.non_max_suppression = inference, non_max_suppression
            sys.modules["_y3_reference"] = ref
            return ref
        except ImportError:
            continue
        finally:
            sys.path[:] = saved_path
            for k in [k for k in sys.modules if k == "yolov3" or k.startswith("yolov3.")]:
                sys.modules["_y3_reference" + k[6:]] = sys.modules.pop(k)
            if saved_mod is not None:
                sys.modules["yolov3"] = saved_mod
    return None


def draw_boxes(img, bbox_tlbr, class_prob=None, class_idx=None, class_names=None):
    """Minimal stand-in for the reference's drawing helper (yolov3/inference.py:97-158) when the
    reference package is not installed: green rectangles + class text."""
    import cv2
    for i, (x1, y1, x2, y2) in enumerate(np.asarray(bbox_tlbr).tolist()):
        cv2.rectangle(img, (x1, y1), (x2, y2), color=(0, 255, 0), thickness=2)
        if class_idx is not None:
            c = int(class_idx[i])
            text = class_names[c] if class_names is not None else str(c)
            cv2.putText(img, text, (x1 + 1, y1 + 13), cv2.FONT_HERSHEY_SIMPLEX, 0.45, (255, 255, 255), thickness=1)


def write_mp4(frames, fps, filepath):
    """Frames -> .mp4 (reference: yolov3/__main__.py:13-33)."""
    import cv2
    if not filepath.endswith(".mp4"):
        filepath += ".mp4"
    h, w = frames[0].shape[:2]
    writer = cv2.VideoWriter(filepath, cv2.VideoWriter_fourcc(*"mp4v"), int(fps) or 30, (w, h))
    for frame in frames:
        writer.write(frame)
    writer.release()


def chunks(items, n):
    for i in range(0, len(items), n):
        yield items[i:i + n]


def build_parser():
    p = argparse.ArgumentParser(prog="yolov3")
    src = p.add_argument_group(title="what to run on (exactly one)").add_mutually_exclusive_group(required=True)
    src.add_argument("-C", "--cam", metavar="cam_id", nargs="?", const=0,
                     help="webcam index or capture-stream path (default: device 0)")
    src.add_argument("-I", "--image", type=pathlib.Path, metavar="<path>",
                     help="one image, or a directory whose images are processed in batches")
    src.add_argument("-V", "--video", type=pathlib.Path, metavar="<path>", help="video file whose frames are processed in batches")
    m = p.add_argument_group(title="network and thresholds")
    m.add_argument("-c", "--config", type=pathlib.Path, required=True, metavar="<path>",
                   help="Darknet .cfg describing the network (required)")
    m.add_argument("-d", "--device", type=str, default="cuda", metavar="<device>",
                   help="sm_100 CUDA device, e.g. cuda or cuda:1 (default cuda; there is no CPU path)")
    m.add_argument("-i", "--iou-thresh", type=float, default=0.3, metavar="<iou>",
                   help="boxes overlapping a kept box of their class by more than this IoU are dropped (default 0.3)")
    m.add_argument("-n", "--class-names", type=pathlib.Path, metavar="<path>",
                   help="text file with one class name per line; without it boxes are labelled by class index")
    m.add_argument("-p", "--prob-thresh", type=float, default=0.05, metavar="<prob>",
                   help="keep detections whose class probability is at least this (default 0.05)")
    m.add_argument("-w", "--weights", type=pathlib.Path, required=True, metavar="<path>",
                   help="Darknet .weights file for that network (required)")
    o = p.add_argument_group(title="display and output")
    o.add_argument("-o", "--output", type=pathlib.Path, metavar="<path>",
                   help="write the annotated frames to this .mp4")
    o.add_argument("--show-fps", action="store_true", help="overlay the processing rate on the webcam view")
    o.add_argument("-v", "--verbose", action="store_true", help="print the device name and timing")
    b = p.add_argument_group(title="batching (this implementation)")
    b.add_argument("-b", "--batch-size", type=int, default=16, metavar="<n>",
                   help="Images / frames per GPU batch in --image and --video modes. [Default 16]")
    b.add_argument("--no-display", action="store_true", help="Do not open windows (headless box).")
    b.add_argument("--save-json", type=pathlib.Path, metavar="<path>",
                   help="Write detections (per image: tlbr boxes, probabilities, class indices) as JSON.")
    return p


def main(argv=None):
    import cv2
    import torch
    args = vars(build_parser().parse_args(argv))
    for k in ("class_names", "config", "weights", "image", "video", "output", "save_json"):
        if args[k] is not None:
            args[k] = str(args[k].expanduser().absolute())
    device = args["device"]
    net = Darknet(args["config"], device=device)
    net.load_weights(args["weights"])
    net.eval()
    if args["verbose"]:
        print(f"Running model on {torch.cuda.get_device_name(torch.device(device))} (yolov3_b200)")
    class_names = None
    if args["class_names"] is not None and os.path.isfile(args["class_names"]):
        with open(args["class_names"], "r") as f:
            class_names = [line.strip() for line in f.readlines()]
    ref = reference_package()
    draw = ref.draw_boxes if ref is not None else draw_boxes
    kw = dict(device=device, prob_thresh=args["prob_thresh"], nms_iou_thresh=args["iou_thresh"])
    show = not args["no_display"]
    dump = {}

    if args["image"]:
        if os.path.isdir(args["image"]):
            image_dir, fnames = args["image"], sorted(os.listdir(args["image"]))
        else:
            image_dir, fname = os.path.split(args["image"])
            fnames = [fname]
        named = [(f, cv2.imread(os.path.join(image_dir, f))) for f in fnames]
        named = [(f, im) for f, im in named if im is not None]
        t0 = time.time()
        batches = list(chunks(named, max(1, args["batch_size"])))
        gen = inference_batches(net, ([im for _, im in b] for b in batches), **kw)  # resize=True, like the CLI
        n = 0
        for b, results in zip(batches, gen):
            for (fname, image), (bbox_tlbr, class_prob, class_idx) in zip(b, results):
                n += 1
                dump[fname] = {"bbox_tlbr": bbox_tlbr.tolist(), "class_prob": class_prob.tolist(),
                               "class_idx": class_idx.tolist()}
                if show:
                    draw(image, bbox_tlbr, class_idx=class_idx, class_names=class_names)
                    cv2.imshow("YOLOv3", image)
                    cv2.waitKey(0)
        if args["verbose"]:
            print(f"{n} images in {time.time() - t0:.3f} s")
    elif args["video"]:
        cap = cv2.VideoCapture(args["video"])
        fps = cap.get(cv2.CAP_PROP_FPS)
        frames_out = [] if args["output"] else None

        def frame_batches():
            while True:
                batch = []
                while len(batch) < max(1, args["batch_size"]):
                    grabbed, frame = cap.read()
                    if not grabbed:
                        break
                    batch.append(frame)
                if not batch:
                    return
                held.append(batch)
                yield batch

        held, idx, stop = [], 0, False
        for results in inference_batches(net, frame_batches(), **kw):
            batch = held.pop(0)
            for frame, (bbox_tlbr, class_prob, class_idx) in zip(batch, results):
                draw(frame, bbox_tlbr, class_idx=class_idx, class_names=class_names)
                dump[f"frame{idx:06d}"] = {"bbox_tlbr": bbox_tlbr.tolist(), "class_prob": class_prob.tolist(),
                                           "class_idx": class_idx.tolist()}
                idx += 1
                if frames_out is not None:
                    frames_out.append(frame)
                if show:
                    cv2.imshow("YOLOv3", frame)
                    stop = stop or cv2.waitKey(1) == ord("q")
            if stop:
                break
        cap.release()
        if frames_out:
            write_mp4(frames_out, fps, args["output"])
    else:  # webcam: latency-bound, one frame per call; the reference's threaded reader / display drive it
        if ref is None:
            raise SystemExit("--cam needs the reference package's VideoGetter/VideoShower (install pytorch-yolov3 "
                             "or keep baseline/_ref): display code is not part of yolov3_b200")
        cam = int(args["cam"]) if isinstance(args["cam"], str) and args["cam"].isdigit() else args["cam"]
        frames = [] if args["output"] else None
        start = time.time()
        try:
            ref.detect_in_cam(net, cam_id=cam, class_names=class_names, show_fps=args["show_fps"], frames=frames, **kw)
        finally:
            if args["output"] and frames:
                write_mp4(frames, 1 / ((time.time() - start) / len(frames)), args["output"])
    if args["save_json"]:
        with open(args["save_json"], "w") as f:
            json.dump(dump, f)
    if show:
        cv2.destroyAllWindows()
    return dump
