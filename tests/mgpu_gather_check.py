"""torchrun worker of tests/test_gpu_round2.py::test_nccl_detection_gather_content_two_ranks: every rank
runs `inference_batches` on its own shard with a `DetectionGather`; rank 0 checks that what NCCL
delivered for EVERY rank and batch equals that rank's detections (recomputed locally from the same
seeded images — the kernels are deterministic)."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import yolov3_b200  # noqa: E402
from yolov3_b200 import distributed as ydist  # noqa: E402


class Recording(ydist.DetectionGather):
    def __init__(self):
        super().__init__()
        self.snaps = []

    def gather_payload(self, eng, total):
        per_rank, counts = super().gather_payload(eng, total)
        if per_rank is not None:  # clone on the batch's stream: ordered behind the NCCL gather
            self.snaps.append(([pr.clone() for pr in per_rank], counts.copy()))
        return per_rank, counts


def batches_of(rank, n_batches, B, S):
    rng = np.random.default_rng(500 + rank)
    return [[rng.integers(0, 256, (S, S, 3), dtype=np.uint8) for _ in range(B)] for _ in range(n_batches)]


def main():
    out_path = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from tools.synth_weights import write_synthetic_weights
    cfg = os.path.join(ROOT, "pytorch-yolov3_b200", "models", "yolov3-tiny.cfg")
    wpath = os.path.join(tempfile.gettempdir(), "y3b200_gather_check_tiny.weights")
    if rank == 0:
        write_synthetic_weights(cfg, 416, wpath + ".tmp", seed=1234)
        os.replace(wpath + ".tmp", wpath)
    dist.barrier()
    net = yolov3_b200.Darknet(cfg, device=str(dev)).load_weights(wpath).eval()
    n_batches, B, S = 4, 8, 416
    mine = batches_of(rank, n_batches, B, S)
    g = Recording()
    local = list(yolov3_b200.inference_batches(net, mine, device=str(dev), prob_thresh=0.05, nms_iou_thresh=0.3,
                                               resize=False, gather=g))
    torch.cuda.synchronize()
    ok, dets = True, 0
    if rank == 0:
        assert len(g.snaps) == n_batches
        for r in range(world):
            theirs = batches_of(r, n_batches, B, S)
            for k in range(n_batches):
                want = local[k] if r == 0 else yolov3_b200.inference(net, theirs[k], device=str(dev), prob_thresh=0.05,
                                                                     nms_iou_thresh=0.3, resize=False)
                per_rank, counts = g.snaps[k]
                got = ydist.unpack_results(per_rank[r].cpu().numpy(), counts[r])
                ok = ok and counts[r].tolist() == [len(w[1]) for w in want]
                for gi, wi in zip(got, want):  # records: classes ascending; inference(): set() order -> compare as sets
                    a = sorted(zip(map(tuple, gi[0].tolist()), gi[1].tolist(), gi[2].tolist()))
                    b = sorted(zip(map(tuple, wi[0].tolist()), wi[1].tolist(), wi[2].tolist()))
                    ok = ok and a == b
                    dets += len(b)
        json.dump({"ok": bool(ok), "batches": n_batches, "detections": int(dets), "world": world}, open(out_path, "w"))
    torch.cuda.synchronize()
    del net, g, local
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
