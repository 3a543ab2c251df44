"""Micro-benchmark: the same GEMM-shaped convolution with A loaded through the 2-D tiled tensor
map vs the im2col tensor map, single-CTA vs CTA-pair tiles (run on the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))
import torch  # noqa: E402
from yolov3_b200 import _lib  # noqa: E402


def timed(fn, iters=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def case(name, n, h, cin, cout, k, **kw):
    dev = torch.device("cuda:0")
    x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, k, k, cin, device=dev).to(torch.bfloat16)
    b = torch.zeros(cout, device=dev)
    y = torch.empty(n, h, h, cout, device=dev, dtype=torch.bfloat16)
    flops = 2 * n * h * h * cout * cin * k * k

    def fn():
        _lib.conv2d(x.data_ptr(), w, b, y.data_ptr(), n=n, h=h, w_in=h, cin=cin, cout=cout, ksize=k, stride=1,
                    pad=(k - 1) // 2, ld_x=cin, ld_y=cout, leaky=True, **kw)
    t = timed(fn)
    print(f"{name:40s} {t*1e6:9.1f} us  {flops/t/1e12:8.1f} TFLOP/s")


if __name__ == "__main__":
    for (n, h) in ((64, 52), (64, 26)):
        print(f"--- n={n} h={h}")
        case("1x1 K=1024 N=256 tiled   1cta", n, h, 1024, 256, 1, force_1cta=True)
        case("1x1 K=1024 N=256 im2col  1cta", n, h, 1024, 256, 1, force_1cta=True, force_im2col=True)
        case("1x1 K=1024 N=256 tiled   pair", n, h, 1024, 256, 1)
        case("1x1 K=1024 N=256 im2col  pair", n, h, 1024, 256, 1, force_im2col=True)
        case("3x3 K=1152 N=256 im2col  1cta", n, h, 128, 256, 3, force_1cta=True)
        case("3x3 K=1152 N=256 im2col  pair", n, h, 128, 256, 3)
        case("3x3 K=2304 N=512 im2col  pair", n, h, 256, 512, 3)
        case("1x1 K=1024 N=128 tiled   1cta", n, h, 1024, 128, 1)
        case("1x1 K=1024 N=128 im2col  1cta", n, h, 1024, 128, 1, force_im2col=True)
