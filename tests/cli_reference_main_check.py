"""Subprocess worker of tests/test_gpu_round2.py::test_cli_*: runs the REFERENCE's own, unmodified
`yolov3.__main__.main()` (from baseline/_ref) in image mode with its hot path re-bound to yolov3_b200
exactly as INTEGRATION.md recipe (b) shows; cv2's window calls are stubbed (headless box) and every
`draw_boxes` call is recorded.  argv: <image_dir> <cfg> <weights> <out.json>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
image_dir, cfg, weights, out = sys.argv[1:5]
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))     # the reference package, as installed
sys.path.insert(1, os.path.join(ROOT, "pytorch-yolov3_b200"))  # yolov3_b200 (its `yolov3` alias is shadowed)

import numpy as np  # noqa: E402
np.int = int  # harness shim: the reference uses the alias NumPy removed (yolov3/inference.py:353)
import cv2  # noqa: E402

# ---- INTEGRATION.md recipe (b) -----------------------------------------------------------------
import yolov3  # noqa: E402  (the reference)
import yolov3_b200  # noqa: E402
assert "baseline/_ref" in yolov3.__file__.replace(os.sep, "/"), yolov3.__file__
yolov3.Darknet = yolov3_b200.Darknet
yolov3.inference = yolov3_b200.inference
yolov3.non_max_suppression = yolov3_b200.non_max_suppression
import yolov3.inference as _ri  # noqa: E402
_ri.inference = yolov3_b200.inference
# --------------------------------------------------------------------------------------------------

drawn = []
real_draw = yolov3.draw_boxes


def recording_draw(img, bbox, class_prob=None, class_idx=None, class_names=None):
    drawn.append({"bbox": np.asarray(bbox).tolist(), "cls": np.asarray(class_idx).tolist()})
    return real_draw(img, bbox, class_prob=class_prob, class_idx=class_idx, class_names=class_names)


yolov3.draw_boxes = recording_draw
cv2.imshow = lambda *a, **k: None
cv2.waitKey = lambda *a, **k: 0
cv2.destroyAllWindows = lambda *a, **k: None

from yolov3.__main__ import main  # noqa: E402  (the reference's CLI, unmodified)
sys.argv = ["yolov3", "-I", image_dir, "-c", cfg, "-w", weights, "-d", "cuda:0", "-p", "0.2", "-i", "0.3", "-v"]
main()
json.dump({"files": os.listdir(image_dir), "drawn": drawn}, open(out, "w"))
