"""Does splitting the 64-image batch into independent sub-batch pipelines on their own streams hide
the tile-quantisation tails and the latency-bound decode / NMS kernels?  (run on the GPU box)

    python tools/overlap_probe.py [--batch 64] [--splits 1,2,4] [--iters 30]

Every split S builds S plans of batch/S images, captures each plan's detection program into its own
CUDA graph and replays the S graphs on S streams forked from / joined to one timing stream; device
time by CUDA events, images/s for the whole batch.  Y3_NO_PDL=1 switches programmatic dependent
launch off (a dependent kernel parked in griddepcontrol.wait holds an SM another pipeline could use).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--splits", default="1,2,4")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--alternate", type=int, default=0,
                    help="N > 1: N full-batch plans on N streams, consecutive steps alternate between them "
                         "(step i+1's convolutions can fill the SMs step i's decode / NMS leave idle)")
    args = ap.parse_args()
    import bench
    import yolov3_b200

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    net = yolov3_b200.Darknet(bench.CFG, device="cuda:0").load_weights(bench.weights_file()).eval()
    B = args.batch
    imgs = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (B, 416, 416, 3), dtype=np.uint8)).to(dev)
    key = ("det_u8", bench.PROB_THRESH, bench.IOU_THRESH)
    main_s = torch.cuda.Stream()
    if args.alternate > 1:
        N = args.alternate
        prio = [0] * N
        engs = [net.engine(B, 416, 416, slot=100 + k) for k in range(N)]
        streams = [torch.cuda.Stream(priority=p) for p in prio]
        for e in engs:
            e.in_u8.copy_(imgs)
            e.orig_hw.copy_(torch.tensor([[416, 416]] * B, dtype=torch.int32))
            e.launch(key)
        torch.cuda.synchronize()

        def run(n):
            for st in streams:
                st.wait_stream(main_s)
            for i in range(n):
                with torch.cuda.stream(streams[i % N]):
                    engs[i % N].launch(key)
            for st in streams:
                main_s.wait_stream(st)

        with torch.cuda.stream(main_s):
            run(2 * N)
            main_s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_s)
            run(args.iters)
            e1.record(main_s)
            main_s.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        print(f"alternate {N}: {ms:.3f} ms per {B} images, {B / ms * 1e3:.0f} images/s "
              f"(PDL {'off' if os.environ.get('Y3_NO_PDL') == '1' else 'on'})", flush=True)
        return
    for S in [int(s) for s in args.splits.split(",")]:
        n = B // S
        engs = [net.engine(n, 416, 416, slot=10 * S + k) for k in range(S)]
        streams = [torch.cuda.Stream() for _ in range(S)]
        for k, e in enumerate(engs):
            e.in_u8.copy_(imgs[k * n:(k + 1) * n])
            e.orig_hw.copy_(torch.tensor([[416, 416]] * n, dtype=torch.int32))
            e.launch(key)  # warm-up + capture
        torch.cuda.synchronize()

        def step():
            for e, st in zip(engs, streams):
                st.wait_stream(main_s)
                with torch.cuda.stream(st):
                    e.launch(key)
            for st in streams:
                main_s.wait_stream(st)

        with torch.cuda.stream(main_s):
            for _ in range(3):
                step()
            main_s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_s)
            for _ in range(args.iters):
                step()
            e1.record(main_s)
            main_s.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        kept = sum(int(e.det_counts.sum().item()) for e in engs)
        print(f"splits {S}: {ms:.3f} ms per {B} images, {B / ms * 1e3:.0f} images/s, kept {kept}"
              f"  (PDL {'off' if os.environ.get('Y3_NO_PDL') == '1' else 'on'})", flush=True)


if __name__ == "__main__":
    main()
