"""Per-convolution timing table of one execution plan (run on the GPU box).

    python tools/conv_report.py [--cfg yolov3] [--size 416] [--batch 64] [--out gpurun_out/conv_report.json]

For every conv launch: GEMM shape, CUDA-event time (launch timed alone, 5 iterations), achieved
TFLOP/s, and the two lower bounds it is judged against — tensor (algorithmic FLOPs / measured bf16
peak) and HBM (input + output + weights + residual bytes, each once / measured copy bandwidth).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="yolov3")
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "conv_report.json"))
    ap.add_argument("--only", default="", help="comma-separated block numbers to time (default: all)")
    a = ap.parse_args()
    import yolov3_b200
    from tools.synth_weights import write_synthetic_weights
    cfg = os.path.join(ROOT, "pytorch-yolov3_b200", "models", a.cfg + ".cfg")
    w = f"/tmp/report_{a.cfg}_{a.size}.weights"
    if not os.path.exists(w):
        write_synthetic_weights(cfg, a.size, w)
    net = yolov3_b200.Darknet(cfg, device="cuda:0").load_weights(w).eval()
    eng = net.engine(a.batch, a.size, a.size)
    x = torch.rand(a.batch, 3, a.size, a.size)
    net.forward(x.cuda())
    torch.cuda.synchronize()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    if a.only:
        keep = {int(b) for b in a.only.split(",")}
        eng.conv_ops = [op for op in eng.conv_ops if op[0] in keep]
    total, per = eng.time_convs(iters=5)
    rows = []
    for n, (blk, sec, flops) in enumerate(per):
        info = eng.conv_info["stem"] if (blk == 0 and eng.stem is not None) else eng.conv_info[blk]
        if eng.head_fused and info["kind"] == "conv" and blk + 1 < len(net.blocks) and net.blocks[blk + 1]["type"] == "yolo":
            info = dict(info, kind="head conv + decode", bytes=info["bytes"] - info["M"] * info["N"] * 4)  # no logits written
        t_tensor = flops / (peaks["bf16_tflops"] * 1e12)
        t_hbm = info["bytes"] / (peaks["hbm_gbs"] * 1e9)
        rows.append({"block": blk, "kind": info["kind"], "M": info["M"], "N": info["N"], "K": info["K"], "k": info["k"],
                     "s": info["s"], "ms": sec * 1e3, "tflops": flops / sec / 1e12, "bound_ms": max(t_tensor, t_hbm) * 1e3,
                     "bound": "tensor" if t_tensor > t_hbm else "hbm", "eff": max(t_tensor, t_hbm) / sec})
    out = {"cfg": a.cfg, "size": a.size, "batch": a.batch, "conv_ms": total * 1e3,
           "tflops": eng.conv_flops / total / 1e12, "bound_ms": sum(r["bound_ms"] for r in rows), "rows": rows}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(f"{a.cfg}@{a.size} B={a.batch}: conv total {total*1e3:.3f} ms, {out['tflops']:.1f} TFLOP/s, "
          f"sum of per-layer bounds {out['bound_ms']:.3f} ms")
    print(f"{'blk':>4} {'M':>9} {'N':>5} {'K':>5} k/s {'ms':>8} {'TF/s':>7} {'bound':>6} {'bnd_ms':>7} {'eff':>5}")
    for r in rows:
        print(f"{r['block']:>4} {r['M']:>9} {r['N']:>5} {r['K']:>5} {r['k']}/{r['s']} {r['ms']:>8.4f} {r['tflops']:>7.1f} "
              f"{r['bound']:>6} {r['bound_ms']:>7.4f} {r['eff']:>5.2f}  {r['kind'] if r['kind'] != 'conv' else ''}")


if __name__ == "__main__":
    main()
