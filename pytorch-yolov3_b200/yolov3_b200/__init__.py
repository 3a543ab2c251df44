"""yolov3_b200 — B200-native (sm_100a) implementation of the inference hot path of
nrsyed/pytorch-yolov3 behind the reference's package API.

    import yolov3_b200 as yolov3
    net = yolov3.Darknet("models/yolov3.cfg", device="cuda").load_weights("yolov3.weights").eval()
    results = yolov3.inference(net, images, device="cuda", prob_thresh=0.05, nms_iou_thresh=0.3)

Exports the hot-path subset of the reference's ``yolov3.__all__`` (yolov3/__init__.py:8-12):
``Darknet``, ``inference``, ``non_max_suppression``, ``cxywh_to_tlbr``, plus ``inference_batches`` (the
batched, pipelined loop the reference's CLI leaves as a TODO).  Display / video / COCO
helpers are out of scope and keep coming from the reference package; see
INTEGRATION.md for how its CLI binds to this module.
"""
from .darknet import Darknet, DummyLayer, MaxPool2d, YOLOLayer, blocks2modules, parse_config
from .inference import cxywh_to_tlbr, inference, inference_batches, non_max_suppression, pinned_images, unpin_images

__all__ = ["Darknet", "cxywh_to_tlbr", "non_max_suppression", "inference", "inference_batches", "pinned_images", "unpin_images"]
__version__ = "0.1.0"
