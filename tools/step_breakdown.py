"""Where one benchmark step goes (run on the GPU box): the det_u8 program split into its phases,
each captured alone into a CUDA graph and timed with CUDA events (10 replays), plus the whole step.

    python tools/step_breakdown.py [--batch 64]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

import torch  # noqa: E402


def timed(fn, stream, iters=10):
    with torch.cuda.stream(stream):
        fn()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fn()
        for _ in range(2):
            g.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            g.replay()
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    import bench
    import yolov3_b200
    from yolov3_b200 import _lib
    net = yolov3_b200.Darknet(bench.CFG, device="cuda:0").load_weights(bench.weights_file()).eval()
    B = a.batch
    eng = net.engine(B, bench.SIZE, bench.SIZE)
    eng.in_u8.copy_(torch.from_numpy(bench.synth_images(B, 1234)).cuda())
    eng.orig_hw.copy_(torch.tensor([[bench.SIZE, bench.SIZE]] * B, dtype=torch.int32))
    eng.detect(bench.PROB_THRESH, bench.IOU_THRESH)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()

    def decode():
        eng.counts.zero_()
        for d, logits in eng.head_descs:
            _lib.yolo_decode_cands(d, logits, bench.PROB_THRESH, eng.orig_hw, eng.cands, eng.counts, eng.cap)

    def nms():
        _lib.nms(eng.cands, eng.counts, eng.B, eng.cap, eng.num_classes, bench.IOU_THRESH, 1, eng.sorted, eng.keep,
                 eng.first_box, eng.nms_ws, class_start=eng.class_start, class_kept=eng.class_kept,
                 dev_thresholds=eng.thresh)

    def compact():
        _lib.compact_kept(eng.sorted, eng.keep, eng.counts, eng.B, eng.cap, eng.dets, eng.det_counts, 1)

    rows = [("backbone (stem + convs)", lambda: eng.run_backbone(fused_stem=True)), ("decode x3 + zero", decode),
            ("nms", nms), ("compact", compact),
            ("backbone + decode", lambda: (eng.run_backbone(fused_stem=True), decode())),
            ("backbone + decode + nms", lambda: (eng.run_backbone(fused_stem=True), decode(), nms())),
            ("whole step, stand-alone decode", lambda: (eng.run_backbone(fused_stem=True), eng._detect_tail(bench.PROB_THRESH, bench.IOU_THRESH))),
            ("whole step (fused head decode)", lambda: eng._detect(bench.PROB_THRESH, bench.IOU_THRESH, fused_stem=True))]
    for name, fn in rows:
        print(f"{name:28s} {timed(fn, s):8.3f} ms")
    print("candidates", int(eng.counts.sum()), "kept", int(eng.det_counts.sum()))


if __name__ == "__main__":
    main()
