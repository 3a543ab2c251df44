#!/usr/bin/env python
"""bench.py — YOLOv3-416 images/sec (forward + decode + NMS) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): models/yolov3.cfg fed 416x416, bf16, batch 64 per GPU,
synthetic uint8 images, calibrated random-init weights (tools/synth_weights.py).  A step is one
pass of the hot path over one batch: uint8 BGR -> bf16 NHWC packing, the Darknet forward
(tcgen05 implicit-GEMM convolutions), YOLO decode + threshold, per-class NMS, compaction.

  value : whole-job images/s with the batch already resident in HBM (CUDA-graph replay; device
          time by CUDA events; max over ranks).  Four distinct input batches are rotated and one
          step streams ~5 GB of activations, so nothing survives in the 126 MB L2 between steps.
          Two execution plans (own buffers, own stream) take the steps alternately, so step i+1's
          convolutions fill the SMs that step i's latency-bound decode / NMS tail leaves idle;
          every step still runs the whole path on its own batch (--plans 1: one plan, no overlap).
  e2e   : the same metric through the public API `yolov3_b200.inference()` with HOST images:
          pinned H2D of the batch and D2H of the kept detections inside the timed region.
  roofline : tensor-core bound; achieved = algorithmic conv FLOPs per step / summed CUDA-event
          time of the conv launches of one step (each launch timed alone, eagerly).
  cpu_baseline : the oracle port of the reference's path (torch CPU fp32 forward + NumPy
          post-processing) on this box's host cores, on a bounded sample.
  --impl reference : the reference's CPU implementation of the path (oracle port; the reference is
          pure Python/torch, nothing to compile) timed on the host cores, same metric and config.

Multi-GPU (torchrun, one rank per GPU): images are independent, so ranks run disjoint batches
(weak scaling, 64 images per GPU); NCCL only gathers detection counts / detections.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

CFG = os.path.join(ROOT, "pytorch-yolov3_b200", "models", "yolov3.cfg")
SIZE = 416
PROB_THRESH, IOU_THRESH = 0.05, 0.3
FLOPS_PER_IMAGE = 65.864075264e9  # sum over convs of 2*Ho*Wo*Cout*Cin*k*k, no padding credit


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm": d["hbm_gbs"], "which": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "which": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_conv_traffic():
    """DRAM bytes moved by the convolution launches of one step, from the committed ncu launch list
    (profiles/*_launches_step_traffic.json, written by tools/summarize_profiles.py); None if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_launches_step_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return d["dram_bytes_per_step"], os.path.relpath(files[-1], ROOT)


def weights_file(tag="yolov3_416"):
    """Seeded calibrated weights, written once per box (rank 0) and shared by both arms."""
    from tools.synth_weights import write_synthetic_weights
    path = os.path.join(tempfile.gettempdir(), f"y3b200_{tag}_seed1234.weights")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        write_synthetic_weights(CFG, SIZE, tmp, seed=1234)
        os.replace(tmp, path)
    return path


def synth_images(batch, seed):
    return np.random.default_rng(seed).integers(0, 256, (batch, SIZE, SIZE, 3), dtype=np.uint8)


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path
# ----------------------------------------------------------------------------------------------
def cpu_reference_step(model, images):
    """One pass of the reference path on the host: preprocess, forward, decode, threshold, NMS."""
    from oracle import darknet_oracle as DO  # cpu_baseline / --impl reference leg only
    from oracle import postprocess_oracle as PO  # cpu_baseline / --impl reference leg only
    blocks, net_info, params = model
    with torch.no_grad():
        out = DO.forward(torch.from_numpy(PO.preprocess(list(images))), blocks, net_info, params)
    return PO.postprocess(out["bbox_xywh"].numpy(), out["class_prob"].numpy(), out["class_idx"].numpy(),
                          [im.shape for im in images], PROB_THRESH, IOU_THRESH)


def load_cpu_reference():
    from oracle import darknet_oracle as DO  # cpu_baseline / --impl reference leg only
    blocks, net_info = DO.load_model(CFG)
    _, params = DO.read_weights(weights_file(), blocks, net_info)
    return blocks, net_info, params


def run_cpu_baseline(n_images=8, min_seconds=10.0):
    """Bounded sample of the workload on the host cores: 8-image sub-batches of the 64-image batch,
    repeated until at least `min_seconds` of CPU work have been timed."""
    model = load_cpu_reference()
    imgs = synth_images(n_images, 4321)
    cpu_reference_step(model, imgs[:1])  # warm-up
    reps = 0
    t0 = time.perf_counter()
    while reps < 2 or time.perf_counter() - t0 < min_seconds:
        cpu_reference_step(model, imgs)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": n_images * reps / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} x {n_images} images of the yolov3-416 workload (torch CPU fp32 forward on "
                      f"{torch.get_num_threads()} threads + NumPy threshold/NMS on 1), {dt:.1f} s"}


def main_reference(args, rank, world):
    if rank != 0:
        return
    n_img = 4
    model = load_cpu_reference()
    imgs = synth_images(n_img, 4321)
    for _ in range(args.warmup):
        cpu_reference_step(model, imgs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(model, imgs)
    dt = time.perf_counter() - t0
    val = n_img * args.steps / dt
    cores = torch.get_num_threads()
    sample = (f"each step = {n_img} images of the workload (bounded sample of the 64-image batch), torch CPU fp32 "
              f"forward on {cores} threads + NumPy threshold/NMS")
    print(json.dumps({
        "impl": "reference", "metric": "YOLOv3-416 images/sec (fwd+decode+NMS)", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "yolov3.cfg 416x416, calibrated random-init weights, prob>=0.05, per-class NMS iou 0.3",
                   "images_per_step": n_img},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def main_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    import yolov3_b200
    from yolov3_b200 import _lib, distributed as ydist

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    if rank == 0:
        wpath = weights_file()
    if world > 1:
        dist.barrier()
    wpath = weights_file()
    net = yolov3_b200.Darknet(CFG, device=str(dev)).load_weights(wpath).eval()
    eng = net.engine(B, SIZE, SIZE)

    # four distinct input batches resident in HBM (133 MB > L2), rotated step by step
    host_batches = [synth_images(B, 1234 + 17 * rank + i) for i in range(4)]
    dev_batches = [torch.from_numpy(b).to(dev) for b in host_batches]
    all_counts = torch.zeros(world * B, dtype=torch.int32, device=dev)
    key = ("det_u8", PROB_THRESH, IOU_THRESH)
    # the steps alternate between `plans` independent execution plans, each on its own stream
    P = max(1, args.plans)
    plans = [eng] if P == 1 else [net.engine(B, SIZE, SIZE, slot=200 + k, concurrent=os.environ.get('Y3_PLANS_PDL', '0') != '1')
                                    for k in range(P)]
    streams = [torch.cuda.Stream(device=dev) for _ in plans]
    for pl in plans:
        pl.orig_hw.copy_(torch.tensor([[SIZE, SIZE]] * B, dtype=torch.int32))
    torch.cuda.synchronize()

    def step(i):
        pl = plans[i % P]
        with torch.cuda.stream(streams[i % P]):
            pl.in_u8.copy_(dev_batches[i % 4], non_blocking=True)
            pl.launch(key)
            if world > 1:  # the path's only collective: detection counts (payload gathered in e2e)
                dist.all_gather_into_tensor(all_counts, pl.det_counts)

    def run(n):
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for i in range(n):
            step(i)
        for st in streams:
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run(args.warmup)
    barrier()
    launches_per_step = plans[0].launches(key)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run(args.steps)
    e1.record()
    barrier()
    dt = e0.elapsed_time(e1) * 1e-3
    if rank == 0:  # clocks are sampled during the device-timed region only: 50 Hz nvidia-smi polling perturbs host-paced code
        clocks = sampler.stop()
    last = plans[(args.steps - 1) % P]
    kept_last = int(last.det_counts.sum().item())
    cands_last = int(last.counts.sum().item())

    # ---- e2e through the public API, host buffers ------------------------------------------------
    host_lists = [list(b) for b in host_batches]
    e2e_steps = max(3, min(args.steps, 100))  # ~0.7 s of host-paced calls: single hiccups (VM scheduling) average out
    for i in range(12):  # the four rotating batches three times: the pinned-memory cache reaches its steady state
        yolov3_b200.inference(net, host_lists[i % 4], device=str(dev), prob_thresh=PROB_THRESH,
                              nms_iou_thresh=IOU_THRESH, resize=False)
        if world > 1:  # NCCL sets its point-to-point channels up on the first gather: not part of a step
            from yolov3_b200.inference import last_device_outputs
            ydist.gather_outputs(*last_device_outputs(net, B, SIZE, SIZE, dev))
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for i in range(e2e_steps):
        res = yolov3_b200.inference(net, host_lists[i % 4], device=str(dev), prob_thresh=PROB_THRESH,
                                    nms_iou_thresh=IOU_THRESH, resize=False)
        if world > 1:  # the path's collective: every rank's detections gathered on rank 0, device to device
            from yolov3_b200.inference import last_device_outputs
            ydist.gather_outputs(*last_device_outputs(net, B, SIZE, SIZE, dev))
        d2h = sum(len(r[1]) for r in res) * 32 + B * 4 + B * eng.num_classes * 4
    barrier()
    dt_e2e = time.perf_counter() - t0

    # ---- max over ranks ------------------------------------------------------------------------------
    times = torch.tensor([dt, dt_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dt, dt_e2e = times.tolist()

    out = None
    if rank == 0:
        peaks = measured_peaks()
        conv_s, per = eng.time_convs(iters=3)
        traffic, traffic_src = ncu_conv_traffic()
        flops_step = eng.conv_flops
        achieved = flops_step / conv_s / 1e12
        share = conv_s / (dt / args.steps)
        out = {
            "metric": "YOLOv3-416 images/sec (fwd+decode+NMS)", "value": world * B * args.steps / dt,
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "yolov3.cfg (Darknet-53) 416x416, batch 64 per GPU, calibrated random-init weights, "
                                   "uint8 images resident in HBM, prob>=0.05, per-class NMS iou 0.3",
                       "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world} (image sharding)",
                       "l2": "4 rotating input batches (133 MB) + ~5 GB of activations streamed per step: "
                             "self-flushing, inputs larger than L2",
                       "candidates_last_step": cands_last, "kept_last_step": kept_last,
                       "cuda_graph": bool(eng.use_graphs),
                       "pipeline": (f"{P} execution plans (own buffers and stream) take the steps alternately: step i+1's "
                                    "convolutions overlap step i's decode/NMS tail; ms_per_step = timed region / steps")
                       if P > 1 else "one plan, steps back to back"},
            "tensor_fraction_of_step": {"value": world * B * args.steps / dt / world * FLOPS_PER_IMAGE /
                                        (peaks["bf16_burst"] * 1e12), "of": f"{peaks['which']} burst bf16 peak"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_burst"], "traffic": traffic,
                         "traffic_note": f"DRAM read+write bytes of the conv launches of one step (ncu, {traffic_src}); "
                                         "algorithmic: 161.3 MB/img activations x 64 + 124 MB weights = 10.4 GB",
                         "of": f"{peaks['which']} burst bf16 (kernel launches timed alone); "
                               f"sustained {peaks['bf16_sustained']}",
                         "kernel": f"conv_umma_kernel + conv_chain_kernel ({len(per)} launches/step cover the 75 "
                                   "convolutional blocks; the uint8 stem and the head decode epilogues included)",
                         "conv_ms_per_step": conv_s * 1e3,
                         "conv_share_of_step": share, "flops_per_step": flops_step},
            "e2e": {"value": world * B * e2e_steps / dt_e2e, "unit": "images/s",
                    "h2d_bytes_per_step": B * SIZE * SIZE * 3 + B * 8, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "yolov3_b200.inference(net, list_of_uint8_images, resize=False)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        slow = sorted(per, key=lambda p: -p[1])[:5]
        out["roofline"]["slowest_convs"] = [
            {"block": b, "ms": s * 1e3, "tflops": f / s / 1e12} for b, s, f in slow]
    if world > 1:
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = run_cpu_baseline()
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # ~1 s of device time: enough clock samples
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--plans", type=int, default=2, help="execution plans the steps alternate between (1: no overlap)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
