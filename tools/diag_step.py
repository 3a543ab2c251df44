"""Diagnostics of the benchmark workload on the GPU box: NMS segment-size distribution and a host
profile of `inference()` (where the end-to-end time goes)."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    import bench
    import yolov3_b200
    B = 64
    net = yolov3_b200.Darknet(bench.CFG, device="cuda:0").load_weights(bench.weights_file()).eval()
    eng = net.engine(B, bench.SIZE, bench.SIZE)
    imgs = bench.synth_images(B, 1234)
    lists = list(imgs)
    res = yolov3_b200.inference(net, lists, device="cuda:0", prob_thresh=bench.PROB_THRESH,
                                nms_iou_thresh=bench.IOU_THRESH, resize=False)
    torch.cuda.synchronize()
    counts = eng.counts.cpu().numpy()
    cands = eng.cands.cpu().numpy()
    seg = []
    for i in range(B):
        cls = cands[i, :counts[i], 5]
        seg.append(np.bincount(cls, minlength=80))
    seg = np.stack(seg)
    print("candidates/img: mean %.0f max %d" % (counts.mean(), counts.max()))
    print("segment sizes (img x class): mean %.1f  p50 %d  p90 %d  p99 %d  max %d  >256: %d  >512: %d of %d" % (
        seg.mean(), np.percentile(seg, 50), np.percentile(seg, 90), np.percentile(seg, 99), seg.max(),
        (seg > 256).sum(), (seg > 512).sum(), seg.size))
    print("per-class totals (sorted desc, top 12):", np.sort(seg.sum(0))[::-1][:12])
    print("kept/img: mean %.0f" % np.mean([len(r[1]) for r in res]))

    for _ in range(3):
        yolov3_b200.inference(net, lists, device="cuda:0", prob_thresh=bench.PROB_THRESH,
                              nms_iou_thresh=bench.IOU_THRESH, resize=False)
    t0 = time.perf_counter()
    for _ in range(5):
        yolov3_b200.inference(net, lists, device="cuda:0", prob_thresh=bench.PROB_THRESH,
                              nms_iou_thresh=bench.IOU_THRESH, resize=False)
    print("inference(): %.2f ms per 64-image call" % ((time.perf_counter() - t0) / 5 * 1e3))
    if getattr(net, "_last_trace", None):  # Y3_TRACE=1: host timeline of the last call
        print("host timeline of the last call (ms since entry):")
        for label, t in net._last_trace:
            print(f"  {t * 1e3:7.3f}  {label}")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        yolov3_b200.inference(net, lists, device="cuda:0", prob_thresh=bench.PROB_THRESH,
                              nms_iou_thresh=bench.IOU_THRESH, resize=False)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(s.getvalue())


if __name__ == "__main__":
    main()
