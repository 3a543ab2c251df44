#!/usr/bin/env python
"""bench.py — images/sec of the hot path (forward + decode + NMS) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config yolov3_416|spp_608|tiny_416|nms_stress] [--no-extras]

Workloads (BASELINE.json `configs`; --config picks one, default = configs[1], the headline):
  yolov3_416  models/yolov3.cfg fed 416x416, bf16, batch 64 per GPU            (configs[1], [4])
  spp_608     models/yolov3-spp.cfg 608x608, bf16, batch 32 (SPP max-pools)    (configs[2])
  tiny_416    models/yolov3-tiny.cfg 416x416, batch 64 (configs[0]'s network on the GPU)
  nms_stress  256 images x 10,647 candidates x 80 classes, per-class NMS        (configs[3])
All use synthetic uint8 images and calibrated random-init weights (tools/synth_weights.py).  A step
is one pass of the hot path over one batch: uint8 BGR -> bf16 NHWC packing, the Darknet forward
(tcgen05 implicit-GEMM convolutions), YOLO decode + threshold, per-class NMS, compaction.

  value : whole-job images/s with the batch already resident in HBM (CUDA-graph replay; device time by
          CUDA events; max over ranks) over EXACTLY --steps steps.  The D2H of the detections is NOT in
          `value` (it is in `e2e`).  Four distinct input batches are rotated and one step streams GBs of
          activations, so nothing survives in the 126 MB L2 between steps.  Two execution plans (own
          buffers, own stream) take the steps alternately, so step i+1's convolutions fill the SMs that
          step i's latency-bound decode / NMS tail leaves idle (--plans 1: one plan, no overlap).
  sustained : the same loop continued for >= 200 steps (~1 s), where the 1 kW power cap is engaged —
          `value` at the driver's --steps 20 is a burst number; both are reported with their clocks.
  e2e   : the same metric through the public API with HOST images, `yolov3_b200.inference_batches`
          (the batched loop behind the CLI): H2D of every batch from page-locked frame buffers
          (`yolov3_b200.pinned_images`) and D2H of its kept detections inside the timed region, batches
          pipelined.  `e2e.pageable_inputs` = the same call on ordinary numpy arrays (one extra host pass
          into pinned staging memory); `e2e.sync_call` = one blocking `yolov3_b200.inference()` per batch.
  roofline : tensor-core bound; achieved = algorithmic conv FLOPs per step / summed CUDA-event time of
          the conv launches of one in-order pass over the network (events between launches, so every
          kernel sees the L2 state its real predecessor leaves); `frac_of_step` divides by the whole
          timed step instead (decode/NMS tail and launch gaps included) — the conservative figure.
  hbm_roofline : the memory-bound kernels (pool / SPP / packing / dense decode), each timed alone with
          a cold L2, algorithmic bytes / time against the measured HBM copy bandwidth.
  cpu_baseline : the oracle port of the reference's path (torch CPU fp32 forward + NumPy
          post-processing) on this box's host cores, on a bounded sample.
  cuda_baseline : the reference's algorithm (same torch ops, oracle port) run eagerly on the GPU
          (cuDNN: TF32, fp32, bf16 channels_last) + its host NumPy post-processing — the
          "existing Blackwell kernels" bar.
  --impl reference : the reference's CPU implementation of the path (oracle port; the reference is
          pure Python/torch, nothing to compile) timed on all host cores, same metric and config.

Multi-GPU (torchrun, one rank per GPU): images are independent, so ranks run disjoint batches
(weak scaling); NCCL only all-gathers detection counts / kept records (yolov3_b200.distributed).
"""
import argparse
import faulthandler
import gc
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

MODELS = os.path.join(ROOT, "pytorch-yolov3_b200", "models")
PROB_THRESH, IOU_THRESH = 0.05, 0.3
# FLOPs/img = sum over convs of 2*Ho*Wo*Cout*Cin*k*k, no padding credit
WORKLOADS = {
    "yolov3_416": {"cfg": "yolov3.cfg", "size": 416, "batch": 64, "flops": 65.864075264e9,
                   "metric": "YOLOv3-416 images/sec (fwd+decode+NMS)", "ref_images": 4,
                   "desc": "yolov3.cfg (Darknet-53) 416x416, batch 64 per GPU"},
    "spp_608": {"cfg": "yolov3-spp.cfg", "size": 608, "batch": 32, "flops": 141.449e9,
                "metric": "YOLOv3-SPP-608 images/sec (fwd+decode+NMS)", "ref_images": 2,
                "desc": "yolov3-spp.cfg 608x608, batch 32 per GPU"},
    "tiny_416": {"cfg": "yolov3-tiny.cfg", "size": 416, "batch": 64, "flops": 5.565e9,
                 "metric": "YOLOv3-tiny-416 images/sec (fwd+decode+NMS)", "ref_images": 8,
                 "desc": "yolov3-tiny.cfg 416x416, batch 64 per GPU"},
}
NMS_STRESS = {"images": 256, "boxes": 10647, "classes": 80, "prob_thresh": 0.01, "iou": 0.3,
              "metric": "NMS stress candidates/sec (256 x 10,647 boxes x 80 classes, per-class greedy NMS)"}
CFG, SIZE = os.path.join(MODELS, "yolov3.cfg"), 416  # tools/ import these


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm": d["hbm_gbs"], "which": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "which": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock / throttle reasons / power sampled every ~5 ms through NVML for the life of the run;
    `window(t0, t1)` summarises the samples taken between two host timestamps (the timed region)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4), ("hw_power_brake", 0x80))

    def __init__(self, device):
        self.rows, self.stop_flag, self.thread, self.handle, self.nv = [], False, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device.index or 0)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.handle = None

    def start(self):
        if self.handle is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv, h = self.nv, self.handle
        while not self.stop_flag:
            try:
                self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)

    def window(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        note = None
        if not rows and self.rows:  # region shorter than one sampling period: nearest sample
            mid = 0.5 * (t0 + t1)
            rows = [min(self.rows, key=lambda r: abs(r[0] - mid))]
            note = "timed region shorter than the sampling period: nearest sample"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({name for _, _, bits, _ in rows for name, bit in self.REASONS if bits & bit})
        out = {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": len(rows), "power_w_max": max(r[3] for r in rows)}
        if note:
            out["note"] = note
        return out


def ncu_conv_traffic():
    """DRAM bytes moved by the convolution launches of one step, from the committed ncu launch list
    (profiles/*_launches_step_traffic.json, written by tools/summarize_profiles.py); None if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_launches_step_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return d["dram_bytes_per_step"], os.path.relpath(files[-1], ROOT)


def weights_file(wl_name="yolov3_416"):
    """Seeded calibrated weights, written once per box (rank 0) and shared by both arms."""
    from tools.synth_weights import write_synthetic_weights
    wl = WORKLOADS[wl_name]
    path = os.path.join(tempfile.gettempdir(), f"y3b200_{wl_name}_seed1234.weights")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        write_synthetic_weights(os.path.join(MODELS, wl["cfg"]), wl["size"], tmp, seed=1234)
        os.replace(tmp, path)
    return path


def synth_images(batch, seed, size=416):
    return np.random.default_rng(seed).integers(0, 256, (batch, size, size, 3), dtype=np.uint8)


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


def load_cpu_reference(wl_name):
    from oracle import darknet_oracle as DO  # cpu_baseline / --impl reference leg only
    blocks, net_info = DO.load_model(os.path.join(MODELS, WORKLOADS[wl_name]["cfg"]))
    _, params = DO.read_weights(weights_file(wl_name), blocks, net_info)
    return blocks, net_info, params


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use the whole host."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def run_cpu_baseline(wl_name, n_images=8, min_seconds=10.0):
    """Bounded sample of the workload on the host cores: small sub-batches of the batch, repeated
    until at least `min_seconds` of CPU work have been timed."""
    cores = use_all_host_threads()
    wl = WORKLOADS[wl_name]
    model = load_cpu_reference(wl_name)
    imgs = synth_images(n_images, 4321, wl["size"])
    cpu_reference_step(model, imgs[:1])  # warm-up
    reps = 0
    t0 = time.perf_counter()
    while reps < 2 or time.perf_counter() - t0 < min_seconds:
        cpu_reference_step(model, imgs)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": n_images * reps / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {n_images} images of the {wl_name} workload (torch CPU fp32 forward on "
                      f"{cores} threads + NumPy threshold/NMS on 1), {dt:.1f} s"}


def stress_candidates(seed, n=10647, classes=80, size=416):
    """tie-free synthetic candidates of one image."""
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(0, size, n), rng.uniform(0, size, n)
    w, h = size * (0.02 + 0.4 * rng.uniform(size=n)), size * (0.02 + 0.4 * rng.uniform(size=n))
    box = np.stack([cx, cy, w, h], 1).astype(np.int64)
    tlbr = np.concatenate([box[:, :2] - box[:, 2:] // 2, box[:, :2] + box[:, 2:] // 2], 1)
    cls = rng.integers(0, classes, n).astype(np.int64)
    prob = rng.permutation(np.linspace(0.01, 0.99, n, dtype=np.float64)).astype(np.float32)
    assert len(np.unique(prob)) == n
    return tlbr, prob, cls


def cpu_nms_stress(n_images=4):
    """The reference's NMS (oracle: NumPy restatement, and the compiled C restatement) on a bounded
    sample of the stress workload — cpu_baseline / --impl reference leg only."""
    from oracle import nms_c  # cpu_baseline / --impl reference leg only
    from oracle import postprocess_oracle as PO  # cpu_baseline / --impl reference leg only
    t_np = t_c = 0.0
    for i in range(n_images):
        tlbr, prob, cls = stress_candidates(100 + i)
        t0 = time.perf_counter()
        a = PO.nms(tlbr, prob, cls, NMS_STRESS["iou"])  # cpu_baseline: the reference algorithm in NumPy
        t1 = time.perf_counter()
        b = nms_c.nms(tlbr, prob, cls, NMS_STRESS["iou"])  # cpu_baseline: the same algorithm restated in C
        t2 = time.perf_counter()
        assert a == b
        t_np, t_c = t_np + t1 - t0, t_c + t2 - t1
    return t_np / n_images, t_c / n_images


def main_reference(args, rank, world):
    if rank != 0:
        return
    cores = use_all_host_threads()
    if args.config == "nms_stress":
        per_np, per_c = cpu_nms_stress(4)
        t0 = time.perf_counter()
        n_img = 2
        for s in range(args.steps):
            for i in range(n_img):
                tlbr, prob, cls = stress_candidates(100 + (s * n_img + i) % 8)
                from oracle import postprocess_oracle as PO  # --impl reference leg
                PO.nms(tlbr, prob, cls, NMS_STRESS["iou"])  # reference arm: oracle port of the reference NMS
        dt = time.perf_counter() - t0
        val = n_img * args.steps * NMS_STRESS["boxes"] / dt
        sample = f"each step = {n_img} images x 10,647 candidates through the NumPy port of the reference NMS, 1 thread"
        print(json.dumps({
            "impl": "reference", "metric": NMS_STRESS["metric"], "value": val, "unit": "candidates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
            "config": {"workload": "nms_stress", "images_per_step": n_img, "ms_per_image_numpy": per_np * 1e3,
                       "ms_per_image_c": per_c * 1e3},
            "cpu_baseline": {"value": val, "unit": "candidates/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), flush=True)
        return
    wl = WORKLOADS[args.config]
    n_img = wl["ref_images"]
    model = load_cpu_reference(args.config)
    imgs = synth_images(n_img, 4321, wl["size"])
    for _ in range(args.warmup):
        cpu_reference_step(model, imgs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(model, imgs)
    dt = time.perf_counter() - t0
    val = n_img * args.steps / dt
    sample = (f"each step = {n_img} images of the workload (bounded sample of the {wl['batch']}-image batch), torch CPU "
              f"fp32 forward on {cores} threads + NumPy threshold/NMS")
    print(json.dumps({
        "impl": "reference", "metric": wl["metric"], "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{wl['desc'].split(',')[0]}, calibrated random-init weights, prob>=0.05, per-class NMS "
                               "iou 0.3", "images_per_step": n_img},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ----------------------------------------------------------------------------------------------
# cuda_baseline: the reference's own algorithm on the GPU through stock PyTorch (cuDNN)
# ----------------------------------------------------------------------------------------------
def run_cuda_baseline(wl_name, dev, iters=5):
    from oracle import darknet_oracle as DO  # cuda_baseline leg (a reported baseline, like cpu_baseline)
    from oracle import postprocess_oracle as PO  # cuda_baseline leg
    wl = WORKLOADS[wl_name]
    blocks, net_info, params = load_cpu_reference(wl_name)  # cuda_baseline: same weights as every other leg
    B = wl["batch"]
    imgs = synth_images(B, 1234, wl["size"])
    x32 = torch.from_numpy(PO.preprocess(list(imgs))).to(dev)  # cuda_baseline input, the reference's preprocessing
    out = {"what": "the reference's algorithm (oracle port: the same torch ops in the same order) run eagerly on "
                   "the GPU through stock PyTorch/cuDNN, batch %d; host post-processing = the reference's NumPy "
                   "threshold + NMS (oracle port), timed on a 2-image sample and extrapolated" % B}
    tf32_before = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    last = None
    try:
        for tag, dtype, tf32, cl in (("fp32_tf32", torch.float32, True, False), ("fp32_ieee", torch.float32, False, False),
                                     ("bf16_channels_last", torch.bfloat16, True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            prm = {i: {k: v.to(dev, dtype) for k, v in d.items()} for i, d in params.items()}
            if cl:
                prm = {i: {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v)
                           for k, v in d.items()} for i, d in prm.items()}
            x = x32.to(dtype)
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
            with torch.no_grad():
                for _ in range(2):
                    res = DO.forward(x, blocks, net_info, prm)  # cuda_baseline: reference forward on the GPU
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    res = DO.forward(x, blocks, net_info, prm)  # cuda_baseline
                e1.record()
                torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / iters
            out[tag] = {"forward_ms_per_batch": ms, "forward_images_per_s": B / ms * 1e3}
            if tag == "fp32_tf32":
                last = {k: v.float().cpu().numpy() for k, v in res.items()}
            del prm, x, res
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_before
    t0 = time.perf_counter()
    n_post = 2
    PO.postprocess(last["bbox_xywh"][:n_post], last["class_prob"][:n_post], last["class_idx"][:n_post],  # cuda_baseline
                   [im.shape for im in imgs[:n_post]], PROB_THRESH, IOU_THRESH)
    post = (time.perf_counter() - t0) / n_post
    out["host_postprocess_ms_per_image"] = post * 1e3
    best = max(out[t]["forward_images_per_s"] for t in ("fp32_tf32", "fp32_ieee", "bf16_channels_last"))
    out["images_per_s_forward_only_best"] = best
    out["images_per_s_with_host_nms"] = 1.0 / (1.0 / best + post)
    return out


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def time_cold(fn, dev, flush, iters=5):
    """Median device time of one launch of `fn` with a cold L2 (a 512 MB memset in front, outside the
    event bracket; it also gives the host time to queue the launch behind it)."""
    ts = []
    for _ in range(iters + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.median(ts[1:])


def hbm_roofline(eng, dev, peaks, flush):
    rows = []
    for name, fn, nbytes in eng.memory_bound_ops():
        s = time_cold(fn, dev, flush)
        rows.append({"kernel": name, "algorithmic_bytes": nbytes, "us": s * 1e6, "achieved": nbytes / s / 1e9,
                     "peak": peaks["hbm"], "unit": "GB/s", "frac": nbytes / s / 1e9 / peaks["hbm"]})
    return rows


def bench_network(wl_name, args, dev, rank, world, sampler, full=True):
    """Device-timed loop (+ sustained leg, e2e, rooflines when `full`) of one network workload.
    Returns (dict of this rank's measurements, objects to keep alive until teardown)."""
    import torch.distributed as dist
    import yolov3_b200
    from yolov3_b200 import distributed as ydist

    wl = WORKLOADS[wl_name]
    B, S = (args.batch or wl["batch"]), wl["size"]
    cfg = os.path.join(MODELS, wl["cfg"])
    if rank == 0:
        weights_file(wl_name)
    if world > 1:
        dist.barrier()
    net = yolov3_b200.Darknet(cfg, device=str(dev)).load_weights(weights_file(wl_name)).eval()
    key = ("det_u8", PROB_THRESH, IOU_THRESH)

    # ---- prepare: every plan is built and every CUDA graph captured BEFORE the collective phase ----
    P = max(1, args.plans)
    plans = [net.engine(B, S, S, slot=200 + k, concurrent=(P > 1 and os.environ.get("Y3_PLANS_PDL", "0") != "1"))
             for k in range(P)]
    eng = plans[0]
    streams = [torch.cuda.Stream(device=dev) for _ in plans]
    host_batches = [synth_images(B, 1234 + 17 * rank + i, S) for i in range(4)]
    dev_batches = [torch.from_numpy(b).to(dev) for b in host_batches]
    hw = torch.tensor([[S, S]] * B, dtype=torch.int32)
    for pl in plans:
        pl.orig_hw.copy_(hw)
        pl.in_u8.copy_(dev_batches[0])
        pl.launch(key)
    torch.cuda.synchronize(dev)
    launches_per_step = eng.launches(key)
    host_lists = [list(b) for b in host_batches]
    conv_seq = conv_alone = None
    if full:
        for res in yolov3_b200.inference_batches(net, host_lists[:3], device=str(dev), prob_thresh=PROB_THRESH,
                                                 nms_iou_thresh=IOU_THRESH, resize=False):
            pass  # builds the pipeline's plans and graphs, pins its staging memory
        yolov3_b200.inference(net, host_lists[0], device=str(dev), prob_thresh=PROB_THRESH, nms_iou_thresh=IOU_THRESH,
                              resize=False)
        conv_seq = eng.time_convs_in_sequence(passes=5)
        conv_alone = eng.time_convs(iters=10)
    torch.cuda.synchronize(dev)
    all_counts = torch.zeros(world * B, dtype=torch.int32, device=dev)

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
        torch.cuda.synchronize(dev)

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        run(n)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) * 1e-3, t0, time.perf_counter()

    run(args.warmup)
    dt, t0, t1 = timed(args.steps)
    m = {"B": B, "S": S, "dt": dt, "clocks": sampler.window(t0, t1) if sampler else None,
         "launches_per_step": launches_per_step, "plans": P}
    last = plans[(args.steps - 1) % P]
    m["kept_last"] = int(last.det_counts.sum().item())
    m["cands_last"] = int(last.counts.sum().item())
    if not full:
        return m, (net, plans)

    # ---- sustained leg: >= 200 steps, power cap engaged ------------------------------------------
    n_sus = max(200, args.steps)
    dt_s, t0, t1 = timed(n_sus)
    m["sustained"] = {"steps": n_sus, "dt": dt_s, "clocks": sampler.window(t0, t1) if sampler else None}

    # ---- e2e through the public API, host buffers --------------------------------------------------
    gather = ydist.DetectionGather(dst=None) if world > 1 else None
    e2e_steps = min(max(args.steps, 60), 400)
    # the same four batches in page-locked frame buffers (what a capture / decode pipeline hands over):
    # uploaded by DMA straight from there; `host_lists` (ordinary pageable arrays) need a staging pass first
    pinned_lists = []
    for b in host_batches:
        frames = yolov3_b200.pinned_images(B, S, S)
        frames[...] = b
        pinned_lists.append(list(frames))

    def e2e_loop(n, lists, stats=None):
        kept = 0
        gen = yolov3_b200.inference_batches(net, (lists[i % 4] for i in range(n)), device=str(dev),
                                            prob_thresh=PROB_THRESH, nms_iou_thresh=IOU_THRESH, resize=False,
                                            gather=gather, stats=stats)
        for res in gen:
            kept = sum(len(r[1]) for r in res)
        return kept

    e2e_loop(8, pinned_lists)  # steady state of the pinned-memory cache; NCCL warms up on the first gathers
    barrier()
    t0 = time.perf_counter()
    host_stats = {}
    kept = e2e_loop(e2e_steps, pinned_lists, host_stats)
    barrier()
    m["e2e"] = {"dt": time.perf_counter() - t0, "steps": e2e_steps,
                "host_ms_per_batch": {k: round(v / e2e_steps * 1e3, 3) for k, v in sorted(host_stats.items())},
                "d2h": kept * 44 + eng.meta.numel() * 4, "h2d": B * S * S * 3 + B * 8,
                "gathered_bytes_per_step": (gather.bytes_gathered // (2 * e2e_steps + 16)) if gather else 0}
    e2e_loop(8, host_lists)
    barrier()
    t0 = time.perf_counter()
    host_stats = {}
    e2e_loop(e2e_steps, host_lists, host_stats)
    barrier()
    m["e2e_pageable"] = {"dt": time.perf_counter() - t0, "steps": e2e_steps,
                         "host_ms_per_batch": {k: round(v / e2e_steps * 1e3, 3) for k, v in sorted(host_stats.items())}}
    n_sync = 20
    for i in range(4):
        yolov3_b200.inference(net, host_lists[i % 4], device=str(dev), prob_thresh=PROB_THRESH,
                              nms_iou_thresh=IOU_THRESH, resize=False)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_sync):
        yolov3_b200.inference(net, host_lists[i % 4], device=str(dev), prob_thresh=PROB_THRESH,
                              nms_iou_thresh=IOU_THRESH, resize=False)
    barrier()
    m["e2e_sync"] = {"dt": time.perf_counter() - t0, "steps": n_sync}
    m["conv_seq"], m["conv_alone"] = conv_seq, conv_alone
    m["conv_flops"] = eng.conv_flops
    return m, (net, plans, gather)


def bench_nms_stress(dev, steps, warmup):
    """BASELINE.json configs[3]: y3_nms on 256 images x 10,647 tie-free candidates x 80 classes."""
    from yolov3_b200 import _lib
    N, n, C = NMS_STRESS["images"], NMS_STRESS["boxes"], NMS_STRESS["classes"]
    sets = []
    for s in range(2):  # two candidate sets (2 x 87 MB) alternate: larger than L2 together
        rec = np.zeros((N, n, 8), dtype=np.int32)
        for i in range(N):
            tlbr, prob, cls = stress_candidates(1000 * s + i)
            rec[i, :, 0:4], rec[i, :, 4], rec[i, :, 5], rec[i, :, 6] = tlbr, prob.view(np.int32), cls, np.arange(n)
        sets.append(torch.from_numpy(rec).to(dev))
    counts = torch.full((N,), n, dtype=torch.int32, device=dev)
    sorted_, keep = torch.empty_like(sets[0]), torch.zeros(N, n, dtype=torch.uint8, device=dev)
    first = torch.empty(N, C, dtype=torch.int32, device=dev)
    ws = torch.empty(_lib.nms_workspace_bytes(N, n, C), dtype=torch.uint8, device=dev)

    def step(i):
        _lib.nms(sets[i % 2], counts, N, n, C, NMS_STRESS["iou"], 1, sorted_, keep, first, ws)

    for i in range(max(3, warmup)):
        step(i)
    torch.cuda.synchronize(dev)
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    launches = _lib.launch_count()
    # class-agnostic NMS of the same candidates (non_max_suppression(class_idx=None), inference.py:232-246):
    # one 10,647-box segment per image, the large-segment path of y3_nms; 16 images, timed separately
    NA = 16
    first1 = torch.empty(NA, 1, dtype=torch.int32, device=dev)
    for _ in range(2):
        _lib.nms(sets[0][:NA], counts[:NA], NA, n, 1, NMS_STRESS["iou"], 0, sorted_[:NA], keep[:NA], first1, ws)
    torch.cuda.synchronize(dev)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    _lib.nms(sets[1][:NA], counts[:NA], NA, n, 1, NMS_STRESS["iou"], 0, sorted_[:NA], keep[:NA], first1, ws)
    a1.record()
    torch.cuda.synchronize(dev)
    agnostic_ms = a0.elapsed_time(a1)
    agnostic_kept = int(keep[:NA].sum().item())
    _lib.nms(sets[(steps - 1) % 2], counts, N, n, C, NMS_STRESS["iou"], 1, sorted_, keep, first, ws)  # restore
    torch.cuda.synchronize(dev)
    return {"metric": NMS_STRESS["metric"], "ms_per_step": ms, "value": N * n / ms * 1e3, "unit": "candidates/s",
            "class_agnostic": {"images": NA, "ms_total": agnostic_ms, "ms_per_image_when_batched": agnostic_ms / NA,
                               "kept": agnostic_kept,
                               "note": "one 10,647-box segment per image (one CTA each, 16 images side by side); the "
                                       "reference's NumPy loop needs ~1 s per image (SURVEY.md §6)"},
            "images_per_s": N / ms * 1e3, "kept_last_step": int(keep.sum().item()), "gpu_launches": launches,
            "algorithmic_bytes": N * n * 32 + int(keep.sum().item()) * 4,
            "workload": "256 images x 10,647 candidates (uniform boxes, 80 classes uniform, distinct scores), "
                        "per-class greedy NMS iou 0.3; candidates resident in HBM, two sets alternate"}


def load_parity():
    path = os.path.join(ROOT, "profiles", "r02_parity.json")
    if os.path.exists(path):
        d = json.load(open(path))
        d["source"] = "profiles/r02_parity.json (written by tests/test_gpu_parity.py on the B200, committed)"
        return d
    return None


def assemble(wl_name, m, args, world, peaks):
    """The JSON line of a network workload from rank 0's (max-reduced) measurements."""
    wl = WORKLOADS[wl_name]
    B, dt, steps = m["B"], m["dt"], args.steps
    flops_step = m.get("conv_flops") or B * wl["flops"]
    out = {
        "metric": wl["metric"], "value": world * B * steps / dt, "unit": "images/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{wl['desc']}, calibrated random-init weights, uint8 images resident in HBM, "
                               "prob>=0.05, per-class NMS iou 0.3", "name": wl_name,
                   "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world} (image sharding)",
                   "l2": "4 rotating input batches + GBs of activations streamed per step: self-flushing, inputs "
                         "larger than L2",
                   "candidates_last_step": m["cands_last"], "kept_last_step": m["kept_last"], "cuda_graph": True,
                   "value_excludes": "the D2H of the detections (SURVEY.md §8d counts it; it is inside e2e)",
                   "pipeline": (f"{m['plans']} execution plans (own buffers and stream) take the steps alternately: step "
                                "i+1's convolutions overlap step i's decode/NMS tail; ms_per_step = timed region / steps")
                   if m["plans"] > 1 else "one plan, steps back to back"},
        "gpu_launches": m["launches_per_step"] * steps, "clocks": m["clocks"],
    }
    frac_step = B * steps / dt * wl["flops"] / (peaks["bf16_burst"] * 1e12)
    if "sustained" in m:
        s = m["sustained"]
        out["sustained"] = {"steps": s["steps"], "value": world * B * s["steps"] / s["dt"], "unit": "images/s",
                            "ms_per_step": s["dt"] / s["steps"] * 1e3, "clocks": s["clocks"],
                            "tensor_fraction_of_step": B * s["steps"] / s["dt"] * wl["flops"] / (peaks["bf16_burst"] * 1e12)}
    if m.get("conv_seq"):
        conv_s, per = m["conv_seq"]
        alone_s, _ = m["conv_alone"]
        traffic, traffic_src = ncu_conv_traffic() if wl_name == "yolov3_416" else (None, None)
        achieved = flops_step / conv_s / 1e12
        out["roofline"] = {
            "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_burst"], "traffic": traffic,
            "traffic_note": f"DRAM read+write bytes of the conv launches of one step (ncu, {traffic_src})",
            "of": f"{peaks['which']} burst bf16; sustained {peaks['bf16_sustained']}",
            "kernel": f"conv_umma / conv_patch / conv_chain kernels ({len(per)} launches/step cover the convolutional "
                      "blocks; the uint8 stem and the head decode epilogues included)",
            "how": "CUDA events between the launches of one in-order pass over the network (median of 5 passes)",
            "conv_ms_per_step": conv_s * 1e3, "conv_share_of_step": conv_s / (dt / steps),
            "flops_per_step": flops_step, "frac_of_step": frac_step,
            "frac_of_step_note": "algorithmic conv FLOPs / WHOLE timed step (decode, NMS, launch gaps included)",
            "kernels_alone": {"conv_ms_per_step": alone_s * 1e3, "frac": flops_step / alone_s / 1e12 / peaks["bf16_burst"],
                              "note": "each launch replayed 10x back to back (operands of small layers stay L2-warm)"},
            "slowest_convs": [{"block": b, "ms": s * 1e3, "tflops": f / s / 1e12}
                              for b, s, f in sorted(per, key=lambda p: -p[1])[:5]]}
    else:
        out["roofline"] = {"bound": "tensor", "achieved": frac_step * peaks["bf16_burst"], "peak": peaks["bf16_burst"],
                           "unit": "TFLOP/s", "frac": frac_step, "traffic": None,
                           "how": "algorithmic conv FLOPs / whole timed step"}
    if "e2e" in m:
        e = m["e2e"]
        out["e2e"] = {"value": world * B * e["steps"] / e["dt"], "unit": "images/s", "h2d_bytes_per_step": e["h2d"],
                      "d2h_bytes_per_step": e["d2h"], "steps": e["steps"],
                      "api": "yolov3_b200.inference_batches(net, iterable of lists of uint8 images, resize=False): "
                             "H2D of every batch from page-locked frame buffers (yolov3_b200.pinned_images) + D2H of "
                             "its detections inside the timed region, 3 batches in flight",
                      "pageable_inputs": {"value": world * B * m["e2e_pageable"]["steps"] / m["e2e_pageable"]["dt"],
                                          "unit": "images/s",
                                          "host_ms_per_batch_rank0": m["e2e_pageable"]["host_ms_per_batch"],
                                          "note": "same call, images in ordinary (pageable) numpy arrays: one extra "
                                                  "host pass stacks them into pinned staging memory"},
                      "nccl_gathered_bytes_per_step": e["gathered_bytes_per_step"],
                      "host_ms_per_batch": {"rank0": e["host_ms_per_batch"], "host_cores": os.cpu_count(),
                                            "note": "where rank 0's host time goes per batch: stage = background "
                                                    "staging thread, wait_gpu = blocked on the batch's kernels (the "
                                                    "healthy state), submit = queueing copies / graph / collectives"},
                      "sync_call": {"value": world * B * m["e2e_sync"]["steps"] / m["e2e_sync"]["dt"], "unit": "images/s",
                                    "api": "yolov3_b200.inference(net, list_of_uint8_images, resize=False), one blocking "
                                           "call per batch"}}
    return out


def teardown(world, keep):
    """Orderly shutdown: finish all device work, drop graphs / plans / pinned buffers while CUDA and NCCL
    are alive, then leave the process group together."""
    import torch.distributed as dist
    torch.cuda.synchronize()
    for obj in keep:
        if hasattr(obj, "invalidate"):
            obj.invalidate()
    keep.clear()
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main_ours(args, rank, local_rank, world):
    import torch.distributed as dist

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    sampler = ClockSampler(dev) if rank == 0 else None
    if sampler:
        sampler.start()
    keep = []

    if args.config == "nms_stress":
        r = bench_nms_stress(dev, args.steps, args.warmup)
        if rank == 0:
            per_np, per_c = cpu_nms_stress(4)  # cpu_baseline: oracle on a bounded sample
            out = {"metric": r["metric"], "value": world * r["value"], "unit": r["unit"], "n_gpus": world,
                   "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                   "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64/f64", "data": "synthetic",
                   "config": {"workload": r["workload"], "name": "nms_stress", "kept_last_step": r["kept_last_step"],
                              "images_per_s": world * r["images_per_s"]},
                   "roofline": {"bound": "hbm", "achieved": r["algorithmic_bytes"] / r["ms_per_step"] / 1e6,
                                "peak": peaks["hbm"], "unit": "GB/s",
                                "frac": r["algorithmic_bytes"] / r["ms_per_step"] / 1e6 / peaks["hbm"], "traffic": None,
                                "note": "NMS is latency/ALU-bound pairwise integer work on a few hundred KB per image, "
                                        "not an HBM stream: the figure to read is kernel time and candidates/s"},
                   "cpu_baseline": {"value": NMS_STRESS["boxes"] / per_np, "unit": "candidates/s", "cores": 1, "kind": "port",
                                    "ms_per_image_numpy_port": per_np * 1e3, "ms_per_image_c_port": per_c * 1e3,
                                    "sample": "4 images x 10,647 candidates, NumPy port of the reference NMS (and its C "
                                              "restatement) on one host thread"},
                   "e2e": None, "gpu_launches": r["gpu_launches"], "clocks": None}
            print(json.dumps(out), flush=True)
        teardown(world, keep)
        return

    m, objs = bench_network(args.config, args, dev, rank, world, sampler, full=True)
    keep.extend(objs)
    # ---- max over ranks ------------------------------------------------------------------------------
    times = torch.tensor([m["dt"], m["sustained"]["dt"], m["e2e"]["dt"], m["e2e_sync"]["dt"], m["e2e_pageable"]["dt"]],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    m["dt"], m["sustained"]["dt"], m["e2e"]["dt"], m["e2e_sync"]["dt"], m["e2e_pageable"]["dt"] = times.tolist()

    if rank == 0:
        out = assemble(args.config, m, args, world, peaks)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        out["hbm_roofline"] = hbm_roofline(objs[1][0], dev, peaks, flush)
        out["parity"] = load_parity()
        if world == 1 and not args.no_extras:
            extras = {}
            keep.clear()
            objs[0].invalidate()  # free the headline workload's plans before building the next ones
            objs = None
            gc.collect()
            torch.cuda.empty_cache()
            small = argparse.Namespace(**{**vars(args), "batch": 0, "steps": max(20, min(args.steps, 50))})
            for name in ("spp_608", "tiny_416"):
                if name == args.config:
                    continue
                mm, oo = bench_network(name, small, dev, 0, 1, sampler, full=True)
                line = assemble(name, mm, small, 1, peaks)
                line["hbm_roofline"] = hbm_roofline(oo[1][0], dev, peaks, flush)
                extras[name] = line
                oo[0].invalidate()
                del oo, mm
                gc.collect()
                torch.cuda.empty_cache()
            r = bench_nms_stress(dev, 20, 3)
            per_np, per_c = cpu_nms_stress(2)  # cpu_baseline of the NMS config: oracle, bounded sample
            r["cpu_baseline"] = {"ms_per_image_numpy_port": per_np * 1e3, "ms_per_image_c_port": per_c * 1e3, "cores": 1,
                                 "kind": "port", "sample": "2 images x 10,647 candidates"}
            r["gpu_ms_per_image"] = r["ms_per_step"] / NMS_STRESS["images"]
            extras["nms_stress"] = r
            out["other_configs"] = extras
            try:
                out["cuda_baseline"] = run_cuda_baseline(args.config, dev)
            except Exception as e:  # a baseline leg must never take the bench line down
                out["cuda_baseline"] = {"error": repr(e)[:300]}
        del flush
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = run_cpu_baseline(args.config)
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), flush=True)  # before the teardown: a late failure must not lose the line
    if sampler:
        sampler.stop()
    del objs, m
    teardown(world, keep)


def main():
    faulthandler.enable()  # a fatal signal (SIGABRT from a C++ terminate, SIGSEGV) prints the Python stack
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # ~1 s of device time
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=0, help="images per GPU and step (0: the workload's own)")
    ap.add_argument("--plans", type=int, default=2, help="execution plans the steps alternate between (1: no overlap)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="yolov3_416", choices=list(WORKLOADS) + ["nms_stress"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other configs / cuda_baseline legs at N=1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        main_reference(args, rank, world)
        return
    from yolov3_b200 import _lib
    rec = _lib.enable_trap_record(torch.device("cuda", local_rank))  # where a kernel watchdog trap would be recorded
    try:
        main_ours(args, rank, local_rank, world)
    except BaseException:
        note = _lib.describe_trap_record(rec)
        if note:
            print(f"[rank {rank}] {note}", file=sys.stderr, flush=True)
        raise


if __name__ == "__main__":
    main()
