"""Calibrated random-init Darknet ``.weights`` generator for benchmarks and smoke runs.

No trained weights exist offline (the reference's get_weights.sh needs the network), and
PyTorch's default initialisation makes activations vanish through 75 layers so that no box
ever passes a threshold.  This writes a seeded file in the reference's
``.weights`` format (yolov3/darknet.py:407-476: 5 x int32 header, then per conv block
``[bn bias, bn weight, bn mean, bn var]`` or ``[bias]``, then the OIHW kernel) whose activations
stay O(1): He-style kernels, BN gamma~U(0.8,1.2), beta~N(0,0.1), head bias~N(0,0.5); one
calibration pass sets every BN's running statistics to the batch statistics of its input and
scales each head so its logits have std 2.

This is INPUT GENERATION, not the inference path: the calibration pass uses plain torch ops
(on the GPU when there is one).  Both bench arms load the same file.
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))


def write_synthetic_weights(cfg_path, size, out_path, seed=1234, device=None):
    from yolov3_b200.darknet import parse_config
    blocks, net_info = parse_config(cfg_path)
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(2, net_info["channels"], size, size, generator=g).to(dev)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    outs, chunks = [], [np.asarray([0, 2, 0, 0, 0], dtype=np.int32).tobytes()]
    try:
        with torch.no_grad():
            for i, b in enumerate(blocks):
                t = b["type"]
                if t == "convolutional":
                    cin, cout, k = x.shape[1], b["filters"], b["size"]
                    pad = (k - 1) // 2 if "pad" in b else 0
                    W = (torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (1.01 * cin * k * k))).to(dev)
                    if "batch_normalize" in b:
                        gamma = (torch.rand(cout, generator=g) * 0.4 + 0.8).to(dev)
                        beta = (torch.randn(cout, generator=g) * 0.1).to(dev)
                        y = F.conv2d(x, W, None, stride=b["stride"], padding=pad)
                        mean, var = y.mean(dim=(0, 2, 3)), y.var(dim=(0, 2, 3), unbiased=False)
                        y = (y - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-5) * \
                            gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
                        chunks += [v.float().cpu().numpy().tobytes() for v in (beta, gamma, mean, var)]
                    else:
                        bias = (torch.randn(cout, generator=g) * 0.5).to(dev)
                        y = F.conv2d(x, W, None, stride=b["stride"], padding=pad)
                        W = W * (2.0 / float(y.std()))
                        y = F.conv2d(x, W, bias, stride=b["stride"], padding=pad)
                        chunks.append(bias.float().cpu().numpy().tobytes())
                    chunks.append(W.float().cpu().numpy().tobytes())
                    x = F.leaky_relu(y, 0.1) if b["activation"] == "leaky" else y
                elif t == "maxpool":
                    k, s = b["size"], b["stride"]
                    xp = F.pad(x, (0, k - 1, 0, k - 1)) if (k > 1 and s == 1) else x
                    x = F.max_pool2d(xp, k, s)
                elif t == "upsample":
                    x = F.interpolate(x, scale_factor=b["stride"], mode="nearest")
                elif t == "route":
                    x = torch.cat([outs[j if j >= 0 else i + j] for j in b["layers"]], dim=1)
                elif t == "shortcut":
                    x = outs[i - 1] + outs[i + b["from"]]
                outs.append(x)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    with open(out_path, "wb") as f:
        for c in chunks:
            f.write(c)
    return out_path


if __name__ == "__main__":
    cfg, size, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    print(write_synthetic_weights(cfg, size, out))
