"""Where multi-GPU end-to-end time goes (run under torchrun on the GPU box): inference() alone vs
inference() + device-to-device gather of the detections, per rank."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import bench
    import yolov3_b200
    from yolov3_b200 import distributed as ydist
    from yolov3_b200.inference import last_device_outputs
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        bench.weights_file()
    dist.barrier()
    net = yolov3_b200.Darknet(bench.CFG, device=str(dev)).load_weights(bench.weights_file()).eval()
    imgs = list(bench.synth_images(64, 1234 + rank))
    kw = dict(device=str(dev), prob_thresh=bench.PROB_THRESH, nms_iou_thresh=bench.IOU_THRESH, resize=False)
    for _ in range(4):
        yolov3_b200.inference(net, imgs, **kw)
        ydist.gather_outputs(*last_device_outputs(net, 64, 416, 416, dev))
    for mode in ("inference only", "inference + gather", "inference only, ranks in lockstep"):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tg = 0.0
        for _ in range(20):
            yolov3_b200.inference(net, imgs, **kw)
            if mode == "inference + gather":
                t1 = time.perf_counter()
                ydist.gather_outputs(*last_device_outputs(net, 64, 416, 416, dev))
                tg += time.perf_counter() - t1
            elif mode.endswith("lockstep"):
                dist.barrier()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        print(f"rank {rank} cores {os.cpu_count()} {mode}: {dt * 1e3:.2f} ms/step, gather {tg / 20 * 1e3:.2f} ms", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
