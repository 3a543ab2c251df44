"""End-to-end loop only (inference_batches with host images), under torchrun or alone, with the host-side
time breakdown — to see what limits e2e scaling on a multi-GPU box.

    [torchrun ...] python tools/e2e_probe.py [--batches 60] [--no-gather] [--tag name]
Environment knobs it is meant to be swept over: Y3_SPIN_SYNC=1, Y3_STAGE_NT=0|1, Y3_STAGE_THREADS=n."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))


KEYS = ("stage", "wait_stage", "submit", "wait_gpu", "wait_copy", "build")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=60)
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    import yolov3_b200
    from yolov3_b200 import distributed as ydist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        bench.weights_file()
    if world > 1:
        dist.barrier()
    net = yolov3_b200.Darknet(bench.CFG, device=str(dev)).load_weights(bench.weights_file()).eval()
    lists = [list(bench.synth_images(64, 1234 + 17 * rank + i)) for i in range(4)]
    if os.environ.get("Y3_PROBE_PINNED", "1") == "1":
        pl = []
        for b in lists:
            frames = yolov3_b200.pinned_images(64, 416, 416)
            frames[...] = np.stack(b)
            pl.append(list(frames))
        lists = pl
    gather = ydist.DetectionGather() if (world > 1 and not a.no_gather) else None

    def loop(n, stats=None):
        for _ in yolov3_b200.inference_batches(net, (lists[i % 4] for i in range(n)), device=str(dev), prob_thresh=0.05,
                                               nms_iou_thresh=0.3, resize=False, gather=gather, stats=stats):
            pass

    loop(8)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stats = {}
    t0 = time.perf_counter()
    loop(a.batches, stats)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt] + [stats.get(k, 0.0) for k in KEYS],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        v = t.tolist()
        per = [round(x / a.batches * 1e3, 3) for x in v[1:]]
        print(json.dumps({"tag": a.tag, "world": world, "images_per_s": round(world * 64 * a.batches / v[0]),
                          "ms_per_batch": round(v[0] / a.batches * 1e3, 3),
                          "host_ms_per_batch_max_over_ranks": dict(zip(KEYS, per)),
                          "env": {k: os.environ.get(k) for k in ("Y3_SPIN_SYNC", "Y3_STAGE_NT", "Y3_STAGE_THREADS", "Y3_PROBE_PINNED")}}), flush=True)
    del net, gather
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
