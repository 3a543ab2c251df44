"""One eager (graph-less) step of the benchmark workload, for ncu:

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_umma|decode|nms|im2col|pack|maxpool|spp' \
        --csv --log-file gpurun_out/launches.csv python tools/profile_step.py [--batch 64] [--steps 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))
os.environ["Y3_NO_GRAPH"] = "1"

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--cfg", default="yolov3")
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--dense", action="store_true", help="also run Darknet.forward (float input packing + dense decode)")
    a = ap.parse_args()
    import bench
    import yolov3_b200
    cfg = os.path.join(ROOT, "pytorch-yolov3_b200", "models", a.cfg + ".cfg")
    if a.cfg == "yolov3" and a.size == 416:
        w = bench.weights_file()
    else:
        from tools.synth_weights import write_synthetic_weights
        w = f"/tmp/prof_{a.cfg}_{a.size}.weights"
        if not os.path.exists(w):
            write_synthetic_weights(cfg, a.size, w)
    net = yolov3_b200.Darknet(cfg, device="cuda:0").load_weights(w).eval()
    eng = net.engine(a.batch, a.size, a.size)
    imgs = torch.randint(0, 256, (a.batch, a.size, a.size, 3), dtype=torch.uint8, device="cuda:0")
    eng.in_u8.copy_(imgs)
    eng.orig_hw.copy_(torch.tensor([[a.size, a.size]] * a.batch, dtype=torch.int32))
    for _ in range(a.steps + 1):  # first iteration is the plan's own warm-up
        eng.detect(bench.PROB_THRESH, bench.IOU_THRESH)
    torch.cuda.synchronize()
    print("kept", int(eng.det_counts.sum()))
    if a.dense:
        x = torch.rand(a.batch, 3, a.size, a.size, device="cuda:0")
        for _ in range(2):
            net.forward(x)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
