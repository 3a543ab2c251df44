"""In-kernel timeline of CTA 0 of selected convolution launches (run on the GPU box with
Y3_CONV_TRACE=1): where one persistent CTA spends its time — set-up, first operands, per-tile MMA
and epilogue phases, store drain.  Cycles are SM clocks (clock64).

    Y3_CONV_TRACE=1 python tools/conv_trace.py [--blocks 14,38,63,64] [--batch 64]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", default="14,38,63,64,13")
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    import bench
    import yolov3_b200
    from yolov3_b200 import _lib
    net = yolov3_b200.Darknet(bench.CFG, device="cuda:0").load_weights(bench.weights_file()).eval()
    eng = net.engine(a.batch, 416, 416)
    net.forward(torch.rand(a.batch, 3, 416, 416).cuda())
    torch.cuda.synchronize()
    L = _lib.lib()
    L.y3_debug_conv_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    ops = {blk: fn for blk, fn, _ in eng.conv_ops}
    buf = (ctypes.c_ulonglong * 96)()
    for blk in [int(b) for b in a.blocks.split(",")]:
        info = eng.conv_info[blk]
        for _ in range(3):
            ops[blk]()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops[blk]()
        e1.record()
        torch.cuda.synchronize()
        assert L.y3_debug_conv_trace(buf, 96) == 0
        t = list(buf)
        t0 = t[0]
        rel = lambda s: (t[s] - t0) if t[s] >= t0 else None  # noqa: E731
        print(f"block {blk}: M={info['M']} N={info['N']} K={info['K']}  event time {e0.elapsed_time(e1)*1e3:.1f} us; "
              f"CTA 0 lifetime {rel(81)} clk")
        print(f"  setup {rel(1)}  pdl {rel(2)}  first TMA issued {rel(3)}")
        prev_commit = None
        for i in range(12):
            land, commit = rel(8 + 2 * i), rel(9 + 2 * i)
            es, er, ed, eo = (rel(32 + 4 * i + j) for j in range(4))
            if land is None or commit is None or (prev_commit is not None and commit < prev_commit):
                break
            prev_commit = commit
            print(f"  tile {i}: operands landed {land:>7}  mma issued-all {commit:>7} | epi start {es}  acc ready {er}"
                  f"  drained {ed}  store issued {eo}" + (f"  (drain {ed - er})" if ed and er and ed > er else ""))
        print(f"  stores drained {rel(80)}  exit {rel(81)}")
        # stale slots from this launch must not leak into the next one
        ctypes.memset(buf, 0, ctypes.sizeof(buf))


if __name__ == "__main__":
    main()
