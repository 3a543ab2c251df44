"""Inference entry points — host-side mirror of the hot-path half of ``yolov3/inference.py`` of
nrsyed/pytorch-yolov3 (``inference`` :286-368, ``non_max_suppression`` :220-266,
``cxywh_to_tlbr`` :269-283), with identical argument meaning, threshold semantics
(keep ``prob >= prob_thresh``, suppress ``iou > nms_iou_thresh``) and return structures, plus
``inference_batches`` — the batched, pipelined loop the reference's CLI leaves as a TODO
(``# TODO: batch images``, yolov3/__main__.py:157; frame loops yolov3/inference.py:435-544).

Everything between the uint8 images and the kept detections runs on the GPU in one CUDA-graph
replay: BGR->RGB /255 packing, the Darknet forward, YOLO decode + threshold + pixel scaling +
integer truncation + tl/br conversion, per-class NMS and the final int64 / float32 arrays.  The
host copies the images up (pinned memory), reads back how many detections every image kept and
receives the arrays; class groups follow the reference's ``set(class_idx)`` visiting order.
There is no CPU implementation here.
"""
import os
import queue
import threading
import time
from collections import deque

import numpy as np
import torch

from . import _lib

_TRACE = os.environ.get("Y3_TRACE", "0") == "1"
# tools/hang_stress.py: random host delay (ms) after every sub-batch launch of inference(), to sweep the ways
# the sub-batch graphs can interleave on the GPU
_JITTER_MS = float(os.environ.get("Y3_STRESS_JITTER_MS", "0"))
# host threads that stage a batch into pinned memory: the box's cores shared between the ranks of a node
# (torchrun exports LOCAL_WORLD_SIZE), at most 16 — beyond that the copy is memory-bound
_STAGE_THREADS = int(os.environ.get("Y3_STAGE_THREADS", "0")) or max(
    2, min(16, (os.cpu_count() or 2) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))


def cxywh_to_tlbr(bbox_xywh):
    """``(cx, cy, w, h, ...) -> (x1, y1, x2, y2, ...)`` with ``tl = c - wh // 2`` and
    ``br = c + wh // 2``; extra columns pass through (reference: yolov3/inference.py:269-283).
    Pure index arithmetic on the caller's host array (the fused GPU decode does the same
    conversion on device for ``inference``)."""
    bbox_tlbr = np.copy(bbox_xywh)
    half = bbox_xywh[:, 2:4] // 2
    bbox_tlbr[:, :2] = bbox_xywh[:, :2] - half
    bbox_tlbr[:, 2:4] = bbox_xywh[:, :2] + half
    return bbox_tlbr


_INT32_MAX = np.iinfo(np.int32).max


def _set_order(first_box_row):
    """Classes in the order the reference's ``for class_ in set(class_idx)`` loop visits them
    (yolov3/inference.py:247-250).  A Python set's iteration order depends only on the hashes and
    on the order in which DISTINCT keys were inserted, i.e. on each class's first occurrence in
    candidate order — which is ascending box index, recorded on the device by ``y3_nms``.

    CPython fact used as a fast path (checked by tests/test_host_logic.py against real sets): with
    at least 19 distinct keys, all of them small non-negative ints below 128, the table has at
    least 128 slots, every key sits in slot ``key`` and iteration is ascending."""
    present = np.nonzero(first_box_row != _INT32_MAX)[0]
    if present.size >= 19 and present[-1] < 128:
        return present.astype(np.int64)
    by_first_seen = present[np.argsort(first_box_row[present], kind="stable")]
    return np.fromiter(set(np.int64(c) for c in by_first_seen), dtype=np.int64, count=present.size)


def _destinations(class_kept, first_box):
    """Where every (image, class) group of kept detections goes in the flat, image-after-image
    output: class groups of one image in the reference's ``set()`` visiting order.  Returns
    (dst_off int32 [B,C], kept per image int64 [B])."""
    B, C = class_kept.shape
    kept64 = class_kept.astype(np.int64)
    per_image = kept64.sum(axis=1)
    flat = kept64.ravel()
    dst = (np.cumsum(flat) - flat).reshape(B, C)  # ascending class order everywhere ...
    n_present = (first_box != _INT32_MAX).sum(axis=1)
    ascending_ok = (n_present >= 19) & (C <= 128)
    if not ascending_ok.all():  # ... except where the set order is not ascending
        base = np.cumsum(per_image) - per_image
        for i in np.nonzero(~ascending_ok)[0]:
            order = _set_order(first_box[i])
            k = kept64[i, order]
            dst[i, order] = base[i] + np.cumsum(k) - k
    return dst.astype(np.int32), per_image


def _order_like_reference(cls_sorted, first_box_row):
    """Permutation taking records sorted by (class asc, prob desc) to the reference's output
    order: class groups in ``set()`` order, prob descending inside a group."""
    if cls_sorted.size == 0:
        return np.zeros(0, dtype=np.int64)
    order = _set_order(first_box_row)
    starts = np.searchsorted(cls_sorted, order, side="left")
    lens = np.searchsorted(cls_sorted, order, side="right") - starts
    # concatenation of the ranges [start_k, start_k + len_k) without a Python loop
    offs = np.cumsum(lens) - lens
    return np.arange(int(lens.sum()), dtype=np.int64) + np.repeat(starts - offs, lens)


def _pinned(shape, dtype):
    return torch.empty(shape, dtype=dtype, pin_memory=True)


# page-locked image buffers handed out by pinned_images(): base address -> (tensor [n,H,W,3], bytes per image)
_PINNED_IMAGES = {}


def pinned_images(n, height, width):
    """A uint8 ``[n, height, width, 3]`` host array in page-locked memory, for callers that own their frame
    buffers (a capture ring, a decoder's output pool).  Batches whose images are consecutive views of such
    an array are uploaded by DMA straight from it; any other array is first stacked into a pinned staging
    buffer by the library's host threads (one extra pass over host memory).  Do not overwrite the images
    of a batch before ``inference`` returned / ``inference_batches`` yielded its results."""
    t = torch.empty((n, height, width, 3), dtype=torch.uint8, pin_memory=True)
    a = t.numpy()
    _PINNED_IMAGES[a.ctypes.data] = (t, height * width * 3)
    return a


def unpin_images(array):
    """Forget a ``pinned_images`` array (its page-locked memory is released once the array itself is)."""
    _PINNED_IMAGES.pop(array.ctypes.data, None)


def _pinned_view(images, H, W):
    """The pinned tensor slice ``[B, H, W, 3]`` the batch's images are consecutive views of, or None."""
    if not _PINNED_IMAGES:
        return None
    per = H * W * 3
    p0 = images[0].ctypes.data
    for base, (t, nbytes_img) in _PINNED_IMAGES.items():
        if nbytes_img != per or not (base <= p0 < base + t.numel()) or (p0 - base) % per:
            continue
        first = (p0 - base) // per
        if first + len(images) > t.shape[0]:
            return None
        for i, im in enumerate(images):
            if not im.flags.c_contiguous or im.ctypes.data != p0 + i * per:
                return None
        return t[first:first + len(images)]
    return None


def _stack_into(dst, images):
    """``np.stack(images)`` straight into the pinned staging buffer: plain memcpy's on a few host
    threads of the library (no interpreter lock held); non-contiguous inputs are compacted first."""
    images = [im if im.flags.c_contiguous else np.ascontiguousarray(im) for im in images]
    _lib.stage_images(dst, images, _STAGE_THREADS)


def _prepare(net, images, resize):
    """Argument handling shared by ``inference`` and ``inference_batches`` (reference:
    yolov3/inference.py:314-326): list-wrap, remember the original shapes, optional cv2 resize to the
    cfg's ``[net]`` size, shape / dtype validation.  Returns (images, orig_shapes, B, H, W)."""
    if not isinstance(images, (list, tuple)):
        images = [images]
    images = list(images)
    orig_shapes = [im.shape for im in images]
    if resize:
        import cv2
        net_shape = (net.net_info["height"], net.net_info["width"])
        images = [cv2.resize(im, net_shape) if im.shape[:2] != net_shape else im for im in images]
    if images[0].ndim != 3 or images[0].shape[2] != 3:
        raise ValueError(f"images must be HxWx3 uint8 BGR arrays, got shape {images[0].shape}")
    first = images[0].shape
    for im in images:
        if im.shape != first:  # the reference fails in np.stack with this error type
            raise ValueError("all input arrays must have the same shape")
        if im.dtype != np.uint8:
            raise ValueError(f"images must be uint8, got {im.dtype}")
    return images, orig_shapes, len(images), first[0], first[1]


def _split_meta(m, B, C):
    """Host copy of ``Engine.meta`` -> (per_image [B], total, class_kept [B,C], first_box [B,C])."""
    return m[:B], int(m[B]), m[B + 1:B + 1 + B * C].reshape(B, C), m[B + 1 + B * C:].reshape(B, C)


def _empty_result():
    return [np.zeros((0, 4), np.int64), np.zeros(0, np.float32), np.zeros(0, np.int64)]


def inference(net, images, device="cuda", prob_thresh=0.05, nms_iou_thresh=0.3, resize=True):
    """Run the network on image(s); same contract as the reference (yolov3/inference.py:286-368).

    Args:
        net: ``yolov3_b200.Darknet`` (in eval mode).
        images: one ``HxWx3`` uint8 BGR array or a list of them (one batch).
        device: CUDA device string; must be the device ``net`` runs on.
        prob_thresh: detections with ``class_prob >= prob_thresh`` are kept.
        nms_iou_thresh: per-class NMS suppresses boxes with ``iou > nms_iou_thresh``.
        resize: resize every image to the cfg's ``[net]`` size first (cv2 bilinear, as the
            reference does); otherwise all images must already share one shape.

    Returns:
        list (one entry per image) of ``[bbox_tlbr int64 (K,4), class_prob float32 (K,),
        class_idx int64 (K,)]`` in ORIGINAL-image pixels, unclipped, ordered like the reference
        (class groups in ``set(class_idx)`` order, descending probability inside a group; equal
        probabilities inside a class are visited in ascending box order — the reference's
        ``np.argsort`` leaves that order unspecified).  Coordinates are exact while they fit in int32
        (they saturate beyond, i.e. for ``exp(tw) * anchor`` above ~2e9 pixels).
    """
    t_enter = time.perf_counter()
    dev = _lib.require_device(device)
    net.check_fresh()
    images, orig_shapes, B, H, W = _prepare(net, images, resize)

    # Large batches go through the GPU as a few sub-batches on their own streams and plans: while
    # sub-batch k computes, the host stages and uploads k+1 and finishes k-1 (destinations, emit,
    # download) — the synchronous call hides most of its own host and PCIe time.  Plans that run side
    # by side are captured WITHOUT programmatic dependent launch: with it (+5-8 % on this call) two
    # concurrently replayed graphs of persistent full-SM kernels deadlock once in a few hundred calls —
    # every CTA of a dependent kernel parked in griddepcontrol.wait for a predecessor that never
    # completes (tools/hang_stress.py reproduces it in seconds; a single stream with PDL, and any number
    # of streams without it, run clean).  DESIGN.md §6.
    spans = _sub_batches(B)
    geom = net.geometry(B, H, W)
    io = geom.get("io")
    if io is None:
        multi = len(spans) > 1
        eng0 = net.engine(spans[0][1] - spans[0][0], H, W, slot=1 if multi else 0, concurrent=multi)
        if eng0.device != dev:
            raise RuntimeError(f"net runs on {eng0.device}, inference(device='{device}') requested")
        with torch.cuda.device(dev):
            C, M = eng0.num_classes, eng0.M
            engines = [net.engine(hi - lo, H, W, slot=(k + 1) if multi else 0, concurrent=multi)
                       for k, (lo, hi) in enumerate(spans)]
            io = {"img": _pinned((B, H, W, 3), torch.uint8), "hw": _pinned((B, 2), torch.int32),
                  "meta": [_pinned((e.meta.numel(),), torch.int32) for e in engines],
                  "dst": [_pinned((hi - lo, C), torch.int32) for lo, hi in spans],
                  "engines": engines,
                  "streams": [torch.cuda.Stream(device=dev) for _ in spans] if multi else [None],
                  "out": (torch.empty(B * M, 4, device=dev, dtype=torch.int64),
                          torch.empty(B * M, device=dev, dtype=torch.float32),
                          torch.empty(B * M, device=dev, dtype=torch.int64)) if multi else None}
        geom["io"] = io
    engines = io["engines"]
    if engines[0].device != dev:
        raise RuntimeError(f"net runs on {engines[0].device}, inference(device='{device}') requested")
    key = ("nms_u8", float(prob_thresh), float(nms_iou_thresh))
    io["hw"].numpy()[...] = np.asarray([[s[0], s[1]] for s in orig_shapes], dtype=np.int32)
    results = [None] * B
    trace = [] if _TRACE else None  # (label, seconds since entry): tools/diag_step.py prints it

    def mark(label):
        if trace is not None:
            trace.append((label, time.perf_counter() - t_enter))

    with torch.cuda.device(dev):
        caller = torch.cuda.current_stream()
        streams = [st if st is not None else caller for st in io["streams"]]
        img_np = io["img"].numpy()
        # phase 1: stage + upload + launch, sub-batch after sub-batch
        for (lo, hi), eng, st, meta in zip(spans, engines, streams, io["meta"]):
            _stack_into(img_np[lo:hi], images[lo:hi])
            mark(f"staged {lo}:{hi}")
            if st is not caller:
                st.wait_stream(caller)
            with torch.cuda.stream(st):
                eng.in_u8.copy_(io["img"][lo:hi], non_blocking=True)
                eng.orig_hw.copy_(io["hw"][lo:hi], non_blocking=True)
                eng.launch(key)
                meta.copy_(eng.meta, non_blocking=True)  # kept per (image, class) + first box per class
            mark(f"launched {lo}:{hi}")
            if _JITTER_MS:
                import random
                time.sleep(random.random() * _JITTER_MS * 1e-3)
        # phase 2: per sub-batch, as soon as its NMS is done: destinations -> emit -> download
        base = 0
        pending = []
        for (lo, hi), eng, st, meta, dstbuf in zip(spans, engines, streams, io["meta"], io["dst"]):
            st.synchronize()
            mark(f"nms done {lo}:{hi}")
            _, _, class_kept, first_box = _split_meta(meta.numpy(), eng.B, eng.num_classes)
            dst_off, per_image = _destinations(class_kept, first_box)
            total = int(per_image.sum())
            eng.last_per_image = per_image
            if total == 0:
                for i in range(lo, hi):
                    results[i] = _empty_result()
                continue
            # results live in fresh pinned arrays (torch's caching host allocator recycles them once
            # the caller drops the result), so the device writes the final dtypes and nothing is re-copied
            # (capacities are rounded up to 16 Ki detections, so the allocator sees a handful of distinct sizes and
            # stops calling cudaHostAlloc after the first few batches)
            cap = (total + 16383) // 16384 * 16384
            tlbr, prob, cls = (_pinned((cap, 4), torch.int64)[:total], _pinned((cap,), torch.float32)[:total],
                               _pinned((cap,), torch.int64)[:total])
            with torch.cuda.stream(st):
                dstbuf.numpy()[...] = dst_off + base
                eng.dst_off.copy_(dstbuf, non_blocking=True)
                eng.emit(io["out"])
                o = io["out"] if io["out"] is not None else (eng.out_tlbr, eng.out_prob, eng.out_cls)
                tlbr.copy_(o[0][base:base + total], non_blocking=True)
                prob.copy_(o[1][base:base + total], non_blocking=True)
                cls.copy_(o[2][base:base + total], non_blocking=True)
            pending.append((lo, st, tlbr, prob, cls, per_image))
            base += total
            mark(f"emit queued {lo}:{hi}")
        io["last_total"], io["last_per_image"] = base, np.concatenate([e.last_per_image for e in engines])
        for lo, st, tlbr, prob, cls, per_image in pending:
            st.synchronize()
            tlbr, prob, cls = tlbr.numpy(), prob.numpy(), cls.numpy()
            pos = 0
            for i, k in enumerate(per_image.tolist()):
                results[lo + i] = [tlbr[pos:pos + k], prob[pos:pos + k], cls[pos:pos + k]]
                pos += k
        for st in streams:
            if st is not caller:
                caller.wait_stream(st)
    mark("results built")
    if trace is not None:
        net._last_trace = trace
    return results


def _sub_batches(batch):
    """Contiguous spans ``inference`` pipelines through the GPU.  The upload of the FIRST span and the
    download of the LAST one are the only transfers nothing hides, while small plans use the GPU less
    well, so the spans grow: 1/8, 3/8, 1/2 of the batch from 32 images, two halves from 16, otherwise
    the whole batch.  ``Y3_SUB_BATCHES=n`` forces n equal spans, ``Y3_SUB_SPLIT=a,b,c`` explicit sizes
    (scaled to the batch)."""
    split = os.environ.get("Y3_SUB_SPLIT", "")
    n = int(os.environ.get("Y3_SUB_BATCHES", "0"))
    if split:
        parts = [float(x) for x in split.split(",")]
    elif n:
        parts = [1.0] * max(1, min(n, batch))
    elif batch >= 32:
        parts = [1.0, 3.0, 4.0]
    elif batch >= 16:
        parts = [1.0, 1.0]
    else:
        parts = [1.0]
    total = sum(parts)
    edges, acc = [0], 0.0
    for p_ in parts:
        acc += p_
        edges.append(min(batch, max(edges[-1], int(round(batch * acc / total)))))
    edges[-1] = batch
    return [(lo, hi) for lo, hi in zip(edges[:-1], edges[1:]) if hi > lo]


def last_device_outputs(net, batch, height, width, device):
    """Device-resident final arrays of the last ``inference`` call with this geometry:
    ``(tlbr int64 [K,4], prob float32 [K], cls int64 [K], per_image int64 numpy [B])``."""
    io = net.geometry(batch, height, width)["io"]
    eng = io["engines"][0]
    o = io["out"] if io["out"] is not None else (eng.out_tlbr, eng.out_prob, eng.out_cls)
    k = io["last_total"]
    return o[0][:k], o[1][:k], o[2][:k], io["last_per_image"]


# ----------------------------------------------------------------------------------------------
# batched, pipelined loop (the CLI's image-directory / frame loops)
# ----------------------------------------------------------------------------------------------
class _Slot:
    """One in-flight batch of ``inference_batches``: its own execution plan (buffers + CUDA graph) and
    stream."""

    def __init__(self, net, B, H, W, index, dev):
        # plans that run side by side on different streams are captured without programmatic dependent
        # launch (PDL + concurrent graphs deadlock)
        self.eng = net.engine(B, H, W, slot=100 + index, concurrent=True)
        if self.eng.device != dev:
            raise RuntimeError(f"net runs on {self.eng.device}, device '{dev}' requested")
        self.stream = torch.cuda.Stream(device=dev)
        self.meta = _pinned((self.eng.meta.numel(),), torch.int32)
        # blocking events: the host thread sleeps in synchronize() instead of spinning on a core the staging
        # threads (and, under torchrun, the other ranks) can use.  Y3_SPIN_SYNC=1 restores the spin wait.
        blocking = os.environ.get("Y3_SPIN_SYNC", "0") != "1"
        self.ev_meta = torch.cuda.Event(blocking=blocking)
        self.ev_out = torch.cuda.Event(blocking=blocking)


class _Stager(threading.Thread):
    """Background half of ``inference_batches``: pulls batches from the caller's iterable, validates /
    resizes them and stacks them into pinned buffers (a ring of ``depth + 1`` per geometry, so it runs up
    to two batches ahead of the GPU).  The copy happens inside the library without the interpreter lock,
    so it overlaps the main thread's queueing, waiting and result building."""

    def __init__(self, net, batches, resize, ring, stats, dev):
        super().__init__(daemon=True, name="y3-stager")
        self.net, self.batches, self.resize, self.ring, self.stats, self.dev = net, batches, resize, ring, stats, dev
        self.out = queue.Queue(maxsize=ring)
        self.free = {}      # geometry -> queue of free (img, hw) pinned buffer pairs
        self.stop = False

    def buffers(self, B, H, W):
        q = self.free.get((B, H, W))
        if q is None:
            q = self.free[(B, H, W)] = queue.Queue()
            for _ in range(self.ring):  # [staging image buffer (allocated on first use), original sizes]
                q.put([None, _pinned((B, 2), torch.int32)])
        return q

    def run(self):
        try:
            torch.cuda.set_device(self.dev)  # pinned allocations of this thread belong to the plan's device
            for images in self.batches:
                if self.stop:
                    break
                images, orig_shapes, B, H, W = _prepare(self.net, images, self.resize)
                q = self.buffers(B, H, W)
                bufs = q.get()  # blocks until the GPU has consumed an earlier batch's buffer
                if self.stop:
                    break
                t1 = time.perf_counter()
                src = _pinned_view(images, H, W)  # upload straight from the caller's page-locked images ...
                if src is None:                   # ... or stack them into this entry's pinned staging buffer
                    if bufs[0] is None:
                        bufs[0] = _pinned((B, H, W, 3), torch.uint8)
                    _stack_into(bufs[0].numpy(), images)
                    src = bufs[0]
                bufs[1].numpy()[...] = np.asarray([[s[0], s[1]] for s in orig_shapes], dtype=np.int32)
                if self.stats is not None:
                    self.stats["stage"] = self.stats.get("stage", 0.0) + time.perf_counter() - t1
                self.out.put(("batch", (B, H, W), (src, bufs), q))
            self.out.put(("end", None, None, None))
        except BaseException as e:  # noqa: BLE001 - re-raised in the consumer thread
            self.out.put(("error", e, None, None))


def _reorder_to_set_order(res, class_kept_row, first_box_row):
    """One image's arrays, emitted with class groups ascending, re-ordered to the reference's
    ``set(class_idx)`` group order (needed only when fewer than 19 classes are present, see
    ``_set_order``)."""
    order = _set_order(first_box_row)
    if order.size < 2 or np.all(order[1:] > order[:-1]):
        return res
    kept = class_kept_row.astype(np.int64)
    starts = np.cumsum(kept) - kept
    lens = kept[order]
    offs = np.cumsum(lens) - lens
    perm = np.arange(int(lens.sum()), dtype=np.int64) + np.repeat(starts[order] - offs, lens)
    return [res[0][perm], res[1][perm], res[2][perm]]


def inference_batches(net, batches, device="cuda", prob_thresh=0.05, nms_iou_thresh=0.3, resize=True, depth=3,
                      gather=None, stats=None):
    """``inference`` over a stream of batches, pipelined: a generator that yields, in order, exactly
    what ``inference(net, batch, ...)`` returns for every batch of ``batches`` (an iterable of image
    lists; a bare ndarray counts as a one-image batch).

    While batch k runs on the GPU (one CUDA-graph replay on its own stream and plan), a background thread
    stages batch k+1 / k+2 into pinned memory (the iterable is consumed from that thread) and the calling
    thread uploads k+1 and downloads / hands out batch k-1, so neither PCIe nor host work sits between two
    batches' kernels.  Up to ``depth`` batches are in flight (>= 2; each holds one plan:
    activations + graph).  The final int64 / float32 arrays are written on the device inside the same
    graph (class groups ascending) and land in pinned host arrays; images whose ``set(class_idx)``
    order is not ascending (fewer than 19 distinct classes) are re-ordered on the host.

    ``gather``: optional ``distributed.DetectionGather`` — every batch's kept detections are also
    gathered, device to device, on its destination rank (all ranks must iterate in lock step).
    ``stats``: optional dict; receives the host seconds spent staging images (``stage``, background thread),
    waiting for the stager (``wait_stage``), queueing work (``submit``), waiting for a batch's kernels
    (``wait_gpu``) and for its download (``wait_copy``), and building the result lists (``build``), summed
    over the batches.
    """
    dev = _lib.require_device(device)
    depth = max(2, int(depth))
    pending = deque()
    thr = (float(prob_thresh), float(nms_iou_thresh))
    counters = {}
    program = "gather_u8" if gather is not None else "emit_u8"

    def tick(name, seconds):
        if stats is not None:
            stats[name] = stats.get(name, 0.0) + seconds

    def submit(geom_key, bufs, free_q):
        t1 = time.perf_counter()
        B, H, W = geom_key
        geom = net.geometry(B, H, W)
        slots = geom.setdefault("pipe", [])
        idx = counters.get(geom_key, 0)
        counters[geom_key] = idx + 1
        with torch.cuda.device(dev):
            if len(slots) <= idx % depth:
                slots.append(_Slot(net, B, H, W, len(slots), dev))
            slot = slots[idx % depth]
            while any(it["slot"] is slot for it in pending):  # only when geometries alternate oddly
                yield_ready.append(finish(pending.popleft()))
            eng = slot.eng
            with torch.cuda.stream(slot.stream):
                eng.in_u8.copy_(bufs[0], non_blocking=True)       # bufs = (upload source, ring entry)
                eng.orig_hw.copy_(bufs[1][1], non_blocking=True)
                eng.launch((program,) + thr)
                slot.meta.copy_(eng.meta, non_blocking=True)
                if gather is not None:
                    gather.post_counts(eng)
                slot.ev_meta.record()
        tick("submit", time.perf_counter() - t1)
        return {"slot": slot, "B": B, "stage": 0, "bufs": bufs, "free_q": free_q}

    def stage_a(it):
        """Batch finished on the GPU: read its counts, queue the download of exactly its detections."""
        slot, B = it["slot"], it["B"]
        eng = slot.eng
        t0 = time.perf_counter()
        slot.ev_meta.synchronize()
        t1 = time.perf_counter()
        tick("wait_gpu", t1 - t0)
        it["free_q"].put(it["bufs"][1])  # its upload is long done: the stager may refill the ring entry
        per_image, total, class_kept, first_box = _split_meta(slot.meta.numpy(), B, eng.num_classes)
        it["per_image"] = per_image.copy()
        it["total"] = total
        odd = np.nonzero(((first_box != _INT32_MAX).sum(axis=1) < 19) | (eng.num_classes > 128))[0]
        it["odd"] = [(int(i), class_kept[i].copy(), first_box[i].copy()) for i in odd]
        with torch.cuda.device(dev), torch.cuda.stream(slot.stream):
            if total:
                cap = (total + 16383) // 16384 * 16384
                it["out"] = (_pinned((cap, 4), torch.int64)[:total], _pinned((cap,), torch.float32)[:total],
                             _pinned((cap,), torch.int64)[:total])
                it["out"][0].copy_(eng.out_tlbr[:total], non_blocking=True)
                it["out"][1].copy_(eng.out_prob[:total], non_blocking=True)
                it["out"][2].copy_(eng.out_cls[:total], non_blocking=True)
            if gather is not None:
                gather.gather_payload(eng, total)
            slot.ev_out.record()
        tick("submit", time.perf_counter() - t1)
        it["stage"] = 1

    def finish(it):
        if it["stage"] == 0:
            stage_a(it)
        t0 = time.perf_counter()
        it["slot"].ev_out.synchronize()
        t1 = time.perf_counter()
        tick("wait_copy", t1 - t0)
        if not it["total"]:
            return [_empty_result() for _ in range(it["B"])]
        tlbr, prob, cls = (t.numpy() for t in it["out"])
        results, pos = [], 0
        for k in it["per_image"].tolist():
            results.append([tlbr[pos:pos + k], prob[pos:pos + k], cls[pos:pos + k]])
            pos += k
        for i, ck, fb in it["odd"]:
            results[i] = _reorder_to_set_order(results[i], ck, fb)
        tick("build", time.perf_counter() - t1)
        return results

    net.check_fresh()
    stager = _Stager(net, batches, resize, depth + 1, stats, dev)
    stager.start()
    yield_ready = []
    try:
        while True:
            t0 = time.perf_counter()
            kind, a_, bufs, free_q = stager.out.get()
            tick("wait_stage", time.perf_counter() - t0)
            if kind == "end":
                break
            if kind == "error":
                raise a_
            pending.append(submit(a_, bufs, free_q))
            while yield_ready:
                yield yield_ready.pop(0)
            if len(pending) >= 2 and pending[-2]["stage"] == 0:
                stage_a(pending[-2])
            while len(pending) >= depth:
                yield finish(pending.popleft())
        while pending:
            yield finish(pending.popleft())
    finally:
        stager.stop = True
        for q in list(stager.free.values()):  # wake a stager that waits for a buffer (generator closed early)
            q.put([None, None])
        try:
            while True:  # ... or for room in the hand-over queue
                stager.out.get_nowait()
        except queue.Empty:
            pass
        while pending:  # never leave device work behind that still reads the pinned buffers
            it = pending.popleft()
            it["slot"].stream.synchronize()


def non_max_suppression(bbox_tlbr, class_prob, class_idx=None, iou_thresh=0.3):
    """Greedy NMS on the GPU; returns the kept indices exactly as the reference does
    (yolov3/inference.py:220-266): per class in ``set(class_idx)`` order when ``class_idx`` is
    given (descending probability inside a class), class-agnostic otherwise.

    Args:
        bbox_tlbr: ``Mx4`` integer array (x1, y1, x2, y2); |coordinates| < 2**31 (the reference
            computes in int64; this implementation stores int32 records and raises beyond).
        class_prob: ``M`` probabilities.  Equal probabilities are visited in ascending index order
            (the reference's ``np.argsort(...)[::-1]`` leaves tie order unspecified).
        class_idx: ``M`` class indices in ``[0, 1024)`` or ``None``.
        iou_thresh: boxes with ``iou > iou_thresh`` w.r.t. a kept box are dropped.
    """
    dev = _lib.require_device("cuda")
    bbox_tlbr = np.asarray(bbox_tlbr)
    n = int(bbox_tlbr.shape[0])
    if n == 0:
        return []
    class_prob = np.asarray(class_prob, dtype=np.float32)
    per_class = class_idx is not None
    cls = np.asarray(class_idx, dtype=np.int64) if per_class else np.zeros(n, dtype=np.int64)
    if per_class and (cls.min() < 0 or cls.max() >= 1024):
        raise ValueError("class_idx must lie in [0, 1024)")
    if np.abs(bbox_tlbr[:, :4]).max() >= 2 ** 31:
        raise ValueError("box coordinates must fit in int32")
    num_classes = int(cls.max()) + 1
    rec = np.zeros((1, n, 8), dtype=np.int32)
    rec[0, :, 0:4] = bbox_tlbr[:, :4]
    rec[0, :, 4] = class_prob.view(np.int32)
    rec[0, :, 5] = cls
    rec[0, :, 6] = np.arange(n)
    with torch.cuda.device(dev):
        cands = torch.from_numpy(rec).to(dev)
        counts = torch.tensor([n], dtype=torch.int32, device=dev)
        sorted_ = torch.empty_like(cands)
        keep = torch.zeros(1, n, dtype=torch.uint8, device=dev)
        first = torch.empty(1, num_classes, dtype=torch.int32, device=dev)
        ws = torch.empty(_lib.nms_workspace_bytes(1, n, num_classes), dtype=torch.uint8, device=dev)
        _lib.nms(cands, counts, 1, n, num_classes, iou_thresh, per_class, sorted_, keep, first, ws)
        srt = sorted_[0].cpu().numpy()
        kp = keep[0].cpu().numpy().astype(bool)
        first = first[0].cpu().numpy()
    kept = srt[kp]
    if not per_class:
        return [int(v) for v in kept[:, 6]]
    perm = _order_like_reference(kept[:, 5].astype(np.int64), first)
    return [int(v) for v in kept[perm, 6]]
