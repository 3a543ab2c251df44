"""Multi-GPU plumbing: one process per GPU, image batches sharded by rank, detections gathered.

The hot path has no exchange step (images are independent through forward, decode and NMS), so
the only collective is the gather of kept detections after NMS: an
``all_gather`` of per-image counts followed by an ``all_gather`` of the padded record buffer.
Works with any ``torch.distributed`` backend: NCCL with CUDA tensors on the B200 box, gloo with
CPU tensors in the host-logic tests.
"""
from collections import deque

import numpy as np
import torch
import torch.distributed as dist

from .engine import records_to_numpy


def shard_range(num_images, rank, world_size):
    """Contiguous shard ``[lo, hi)`` of a global batch owned by ``rank`` (remainder spread over
    the first ranks)."""
    base, rem = divmod(num_images, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(results):
    """``inference()`` results of the local images -> (records int32 [K,8], counts int64 [B])."""
    counts = np.asarray([len(r[1]) for r in results], dtype=np.int64)
    rec = np.zeros((int(counts.sum()), 8), dtype=np.int32)
    pos = 0
    for tlbr, prob, cls in results:
        k = len(prob)
        rec[pos:pos + k, 0:4] = tlbr
        rec[pos:pos + k, 4] = np.asarray(prob, dtype=np.float32).view(np.int32)
        rec[pos:pos + k, 5] = cls
        pos += k
    return rec, counts


def unpack_results(rec, counts):
    """Inverse of pack_results."""
    out, pos = [], 0
    for k in counts:
        tlbr, prob, cls, _ = records_to_numpy(rec[pos:pos + int(k)])
        out.append([tlbr, prob, cls])
        pos += int(k)
    return out


def gather_detections(rec, counts, group=None, device=None):
    """All-gather kept detections of every rank.

    Args:
        rec: this rank's records, int32 ``[K, 8]`` (numpy or tensor), images back to back.
        counts: records per local image, ``[B_local]``.
        device: where the collective's tensors live (``cuda:<local_rank>`` for NCCL, ``cpu`` for gloo).
    Returns:
        (records int32 numpy [K_total, 8] in global image order, counts int64 numpy [B_total]);
        ranks may own different numbers of images.
    """
    world = dist.get_world_size(group)
    device = torch.device(device or "cpu")
    rec_t = torch.as_tensor(np.ascontiguousarray(rec) if isinstance(rec, np.ndarray) else rec).to(device, torch.int32)
    cnt_t = torch.as_tensor(np.asarray(counts, dtype=np.int64)).to(device)
    # 1. how many images / records every rank holds
    meta = torch.tensor([cnt_t.numel(), rec_t.shape[0]], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = torch.stack(metas).cpu().numpy()
    max_img, max_rec = int(metas[:, 0].max()), int(metas[:, 1].max())
    # 2. padded counts and records
    cnt_pad = torch.zeros(max(max_img, 1), dtype=torch.int64, device=device)
    cnt_pad[:cnt_t.numel()] = cnt_t
    rec_pad = torch.zeros(max(max_rec, 1), 8, dtype=torch.int32, device=device)
    rec_pad[:rec_t.shape[0]] = rec_t
    all_cnt = [torch.zeros_like(cnt_pad) for _ in range(world)]
    all_rec = [torch.zeros_like(rec_pad) for _ in range(world)]
    dist.all_gather(all_cnt, cnt_pad, group=group)
    dist.all_gather(all_rec, rec_pad, group=group)
    out_cnt = np.concatenate([c.cpu().numpy()[:metas[r, 0]] for r, c in enumerate(all_cnt)])
    out_rec = np.concatenate([t.cpu().numpy()[:metas[r, 1]] for r, t in enumerate(all_rec)])
    return out_rec, out_cnt


def gather_outputs(tlbr, prob, cls, per_image, group=None, dst=0):
    """Gather the final detection arrays of every rank on GLOBAL rank ``dst`` (a member of ``group``)
    after a synchronous ``inference()`` call.  (``inference_batches(..., gather=DetectionGather())`` is
    the pipelined form without host round trips.)

    ``tlbr`` int64 [K,4], ``prob`` float32 [K], ``cls`` int64 [K] are this rank's detections, images
    back to back (``Engine.out_tlbr[:K]`` etc. after ``inference()``), ``per_image`` the detections
    per local image.  The tensors stay where they live (``cuda:<local_rank>`` under NCCL, CPU under
    gloo): one ``all_gather`` of the counts, then one ``gather`` per array, padded to the largest
    rank.  Returns on ``dst``: ``(list of (tlbr, prob, cls) per rank — views of the gathered buffers,
    per-image counts int64 numpy [B_total])``; on the other ranks ``(None, counts)``.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank()  # global, like ``dst`` (dist.gather's dst is a global rank)
    device = tlbr.device
    cnt = torch.as_tensor(np.asarray(per_image, dtype=np.int64)).to(device)
    meta = torch.tensor([cnt.numel(), int(prob.shape[0])], dtype=torch.int64, device=device)
    metas = torch.zeros(world * 2, dtype=torch.int64, device=device)  # flat: gloo accepts no other shape
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.cpu().numpy().reshape(world, 2)
    max_img, max_rec = max(int(metas[:, 0].max()), 1), max(int(metas[:, 1].max()), 1)
    cnt_pad = torch.zeros(max_img, dtype=torch.int64, device=device)
    cnt_pad[:cnt.numel()] = cnt
    all_cnt = torch.zeros(world * max_img, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_cnt, cnt_pad, group=group)
    all_cnt = all_cnt.cpu().numpy().reshape(world, max_img)
    counts = np.concatenate([all_cnt[r, :metas[r, 0]] for r in range(world)])

    def padded(t, shape):
        if t.shape[0] == shape[0]:
            return t.contiguous()
        out = torch.zeros(shape, dtype=t.dtype, device=device)
        out[:t.shape[0]] = t
        return out

    parts = []
    for t, shape in ((tlbr, (max_rec, 4)), (prob, (max_rec,)), (cls, (max_rec,))):
        src = padded(t, shape)
        bufs = [torch.empty_like(src) for _ in range(world)] if rank == dst else None
        dist.gather(src, bufs, dst=dst, group=group)
        parts.append(bufs)
    if rank != dst:
        return None, counts
    per_rank = [(parts[0][r][:metas[r, 1]], parts[1][r][:metas[r, 1]], parts[2][r][:metas[r, 1]]) for r in range(world)]
    return per_rank, counts


class DetectionGather:
    """The hot path's only collective, in the form ``inference_batches`` drives without host round
    trips of its own: per batch ONE small ``all_gather`` of every rank's per-image
    detection counts, issued on the stream right behind the batch's CUDA graph, and — once the host has
    read them together with its own counts (it has to wait for those anyway to size its download) — ONE
    ``all_gather`` of exactly ``max over ranks`` kept records (32-byte ``y3_cand``, rounded up to 4096 of
    them).  Every rank derives the same size from the same gathered counts, so no rank waits for another
    on the host.  (An all-gather rather than a gather onto one rank: over NVSwitch it is a single ring /
    NVLS kernel and one host call, where the grouped send/recv of a rooted gather cost ~2 ms of host time
    per batch at 8 ranks — measured, tools/e2e_probe.py — and made the root's GPU the straggler.)  Works
    with NCCL (device tensors) and gloo (CPU tensors, tests).

    After batch k's payload has been queued, ``last`` = ``(per_rank, counts)`` — on every rank when
    ``dst`` is None, else on global rank ``dst`` only (``(None, counts)`` elsewhere): ``per_rank[r]`` =
    int32 ``[K_r, 8]`` records of rank r's images back to back (per image: class ascending, probability
    descending; ``engine.records_to_numpy`` widens them to the reference's arrays), views of the receive
    buffer (valid once the stream has caught up; overwritten by this slot's next batch), ``counts`` int64
    numpy ``[world, B]``.
    """

    def __init__(self, group=None, dst=0):
        self.group, self.dst = group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank()
        self._queue = deque()
        self.last = None
        self.bytes_gathered = 0

    @staticmethod
    def _host(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=torch.cuda.is_available())

    def post_counts(self, eng):
        """Queue the all-gather of ``eng.det_counts_total`` ([B] per-image counts + total) and its copy
        to the host, on the current stream."""
        st = eng.__dict__.setdefault("_gather_state", {})
        n = eng.det_counts_total.numel()
        if "counts" not in st:
            st["counts"] = torch.zeros(self.world * n, dtype=torch.int32, device=eng.det_counts_total.device)
            st["counts_host"] = self._host((self.world * n,), torch.int32)
        dist.all_gather_into_tensor(st["counts"], eng.det_counts_total, group=self.group)
        st["counts_host"].copy_(st["counts"], non_blocking=True)
        self._queue.append(eng)

    def gather_payload(self, eng, total):
        """Queue the gather of this batch's kept records ``eng.dets`` (call once the copy queued by
        ``post_counts`` has completed, i.e. after the batch's meta event).  ``total`` = this rank's count."""
        assert self._queue.popleft() is eng, "DetectionGather: batches must be finished in submission order"
        st = eng._gather_state
        B = eng.det_counts_total.numel() - 1
        allc = st["counts_host"].numpy().reshape(self.world, B + 1).astype(np.int64)
        assert int(allc[self.rank if self.group is None else dist.get_rank(self.group), B]) == int(total)
        cap = eng.dets.shape[0]
        n = min(cap, max(4096, (int(allc[:, B].max()) + 4095) // 4096 * 4096))
        if "recv" not in st:
            st["recv"] = torch.empty((self.world * cap, 8), dtype=eng.dets.dtype, device=eng.dets.device)
        recv = st["recv"][:self.world * n]
        dist.all_gather_into_tensor(recv.view(-1), eng.dets[:n].reshape(-1), group=self.group)
        self.bytes_gathered += n * 32 * self.world
        counts = allc[:, :B]
        if self.dst is None or self.rank == self.dst:
            self.last = ([recv[r * n:r * n + int(allc[r, B])] for r in range(self.world)], counts)
        else:
            self.last = (None, counts)
        return self.last
