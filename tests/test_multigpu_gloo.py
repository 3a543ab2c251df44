"""CPU, world_size 2, gloo: the N>1 path's host logic — batch sharding and the detection gather
(the only collective of the hot path)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _synth_results(rng, n_img):
    out = []
    for _ in range(n_img):
        k = int(rng.integers(0, 40))
        tlbr = rng.integers(-50, 700, (k, 4)).astype(np.int64)
        prob = rng.random(k, dtype=np.float32)
        cls = rng.integers(0, 80, k).astype(np.int64)
        out.append([tlbr, prob, cls])
    return out


def _worker(rank, world, port, total_images, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from yolov3_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = D.shard_range(total_images, rank, world)
        everything = _synth_results(np.random.default_rng(5), total_images)  # same on every rank
        rec, counts = D.pack_results(everything[lo:hi])
        all_rec, all_counts = D.gather_detections(rec, counts, device="cpu")
        got = D.unpack_results(all_rec, all_counts)
        ok = len(got) == total_images
        for a, b in zip(got, everything):
            ok = ok and all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a, b))
        # device-to-device form (final arrays gathered on rank 0; CPU tensors under gloo)
        import torch
        mine = everything[lo:hi]
        cat = [np.concatenate([m[j] for m in mine]) if mine else np.zeros((0, 4) if j == 0 else 0) for j in range(3)]
        per_rank, cnts = D.gather_outputs(torch.from_numpy(cat[0].astype(np.int64).reshape(-1, 4)),
                                          torch.from_numpy(cat[1].astype(np.float32)),
                                          torch.from_numpy(cat[2].astype(np.int64)), [len(m[1]) for m in mine])
        ok = ok and cnts.tolist() == [len(e[1]) for e in everything]
        if rank == 0:
            for j in range(3):
                whole = np.concatenate([pr[j].numpy() for pr in per_rank])
                ok = ok and np.array_equal(whole, np.concatenate([e[j] for e in everything]))
        else:
            ok = ok and per_rank is None
        q.put((rank, bool(ok), lo, hi))
    finally:
        dist.destroy_process_group()


def _run(total_images, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_gather_detections_world2_even_split():
    res = _run(8, 29611)
    assert [r[1] for r in res] == [True, True]
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 4, 4, 8)


def test_gather_detections_world2_ragged_split():
    res = _run(5, 29612)  # ranks own 3 and 2 images
    assert [r[1] for r in res] == [True, True]
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 3, 3, 5)


def test_shard_range_covers_batch():
    from yolov3_b200.distributed import shard_range
    for n in (1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


class _FakePlan:
    """What DetectionGather reads from an Engine: counts + the three final arrays (CPU tensors here)."""

    def __init__(self, results, cap):
        import torch
        B = len(results)
        per = [len(r[1]) for r in results]
        self.det_counts_total = torch.tensor(per + [sum(per)], dtype=torch.int32)
        self.B = B
        self.dets = torch.zeros(cap, 8, dtype=torch.int32)
        k = sum(per)
        if k:
            from yolov3_b200.distributed import pack_results
            self.dets[:k] = torch.from_numpy(pack_results(results)[0])


def _gather_worker(rank, world, port, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    from yolov3_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        g = D.DetectionGather(dst=0)
        everything = [_synth_results(np.random.default_rng(40 + b), 6) for b in range(3)]  # 3 batches x 6 images
        lo, hi = D.shard_range(6, rank, world)
        plans = [_FakePlan(batch[lo:hi], 20000) for batch in everything]
        # two batches in flight, as inference_batches drives it: counts(k+1) is posted before payload(k)
        g.post_counts(plans[0])
        g.post_counts(plans[1])
        outs = [g.gather_payload(plans[0], int(plans[0].det_counts_total[-1]))]
        g.post_counts(plans[2])
        outs.append(g.gather_payload(plans[1], int(plans[1].det_counts_total[-1])))
        outs.append(g.gather_payload(plans[2], int(plans[2].det_counts_total[-1])))
        for batch, (per_rank, counts) in zip(everything, outs):
            ok = ok and counts.reshape(-1).tolist() == [len(r[1]) for r in batch]
            if rank == 0:
                rec = np.concatenate([pr.numpy() for pr in per_rank])
                got = D.unpack_results(rec, counts.reshape(-1))
                ok = ok and all(np.array_equal(x, y) for g, b in zip(got, batch) for x, y in zip(g, b))
            else:
                ok = ok and per_rank is None
        # gather_outputs on a sub-group whose destination is NOT global rank 0
        if world >= 3:
            sub = dist.new_group([1, 2])
            if rank in (1, 2):
                res = _synth_results(np.random.default_rng(90 + rank), 2)
                cat = [np.concatenate([m[j] for m in res]) for j in range(3)]
                per_rank, cnts = D.gather_outputs(torch.from_numpy(cat[0].reshape(-1, 4)), torch.from_numpy(cat[1]),
                                                  torch.from_numpy(cat[2]), [len(m[1]) for m in res], group=sub, dst=2)
                if rank == 2:
                    exp = [_synth_results(np.random.default_rng(90 + r), 2) for r in (1, 2)]
                    for pr, e in zip(per_rank, exp):
                        ok = ok and np.array_equal(pr[1].numpy(), np.concatenate([m[1] for m in e]))
                else:
                    ok = ok and per_rank is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_detection_gather_pipelined_order_and_subgroup_destination():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 3, 29613, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True), (2, True)]
