"""bf16-matched CPU oracle (TEST INFRASTRUCTURE — not product code).

SURVEY.md §8c / H1-iv: random-weight YOLOv3 amplifies a 1e-6 perturbation 40-90x end to end, so the
fp32 reference and a bf16 pipeline disagree on many detections for reasons that have nothing to do
with correctness.  This module restates the reference's forward (yolov3/darknet.py:351-405) with the
SAME rounding points the B200 plan has, everything else in fp32 on the CPU:

  * BatchNorm folded into the convolution in fp32 (W' = W*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps);
    reference modules: Conv2d/BatchNorm2d built at yolov3/darknet.py:244-257), W' rounded to bf16,
    b' kept fp32;
  * every block output rounded to bf16 (activations are stored as bf16), except YOLO head
    convolutions (fp32 logits) and a convolution whose ONLY consumer is a [shortcut]: the plan adds
    the residual to the fp32 accumulator and rounds once (yolov3/darknet.py:376-379);
  * network input rounded to bf16.

``tests/golden/make_golden.py`` builds the same thing from the reference's OWN modules (deep copy,
fold into conv.weight, BatchNorm turned into a pure bias, forward hooks that round) and asserts the
two agree bit for bit on micro.cfg (golden ``micro_bf16_matched.npz``).
"""
import torch
import torch.nn.functional as F

from . import darknet_oracle as DO


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def fold_bn(prm):
    """(W', b') in fp32 — the fold the engine performs at weight-load time."""
    W = prm["weight"].float()
    if "bn_bias" in prm:
        scale = prm["bn_weight"] / torch.sqrt(prm["bn_var"] + DO.BN_EPS)
        return W * scale.view(-1, 1, 1, 1), prm["bn_bias"] - prm["bn_mean"] * scale
    return W, prm["bias"]


def single_consumer_shortcuts(blocks):
    """{conv block a: shortcut block i} where the shortcut's first operand is conv a and nothing else
    reads a — the residual is then added before the one rounding (engine.py fusion rule)."""
    def root(j):
        while blocks[j]["type"] == "yolo" or (blocks[j]["type"] == "route" and len(blocks[j]["layers"]) == 1):
            j = j - 1 if blocks[j]["type"] == "yolo" else blocks[j]["layers"][0]
        return j
    uses = {}
    for i, b in enumerate(blocks):
        if b["type"] == "route":
            if len(b["layers"]) == 1:
                continue  # alias: its readers are counted instead
            srcs = [root(j) for j in b["layers"]]
        elif b["type"] == "shortcut":
            srcs = [root(i - 1), root(i + b["from"])]
        else:
            srcs = [root(i - 1)] if i > 0 else []
        for s in srcs:
            uses.setdefault(s, []).append(i)
    fused = {}
    for i, b in enumerate(blocks):
        if b["type"] == "shortcut":
            a, r = root(i - 1), root(i + b["from"])
            head = a + 1 < len(blocks) and blocks[a + 1]["type"] == "yolo"
            if blocks[a]["type"] == "convolutional" and uses.get(a) == [i] and a != r and not head:
                fused[a] = i
    return fused


def forward(x, blocks, net_info, params, capture=None):
    """Same contract as darknet_oracle.forward, with the plan's rounding points."""
    fused = single_consumer_shortcuts(blocks)
    keep = DO.resolve_keep(blocks)
    cache, boxes, probs, idxs = {}, [], [], []
    x = bf16(x)
    for i, b in enumerate(blocks):
        t = b["type"]
        if t == "convolutional":
            k, stride, pad, _, leaky = DO.conv_geometry(b)
            W, bias = fold_bn(params[i])
            y = F.conv2d(x, bf16(W), None, stride=stride, padding=pad) + bias.view(1, -1, 1, 1)
            if leaky:
                y = F.leaky_relu(y, 0.1)
            head = i + 1 < len(blocks) and blocks[i + 1]["type"] == "yolo"
            x = y if (head or i in fused) else bf16(y)
        elif t == "maxpool":
            x = DO.maxpool_block(x, b)
        elif t == "upsample":
            x = F.interpolate(x, scale_factor=b["stride"], mode="nearest")
        elif t == "route":
            x = torch.cat([cache[j] for j in b["layers"]], dim=1)
        elif t == "shortcut":
            x = bf16(cache[i - 1] + cache[i + b["from"]])
        elif t == "yolo":
            anchors = [b["anchors"][m] for m in b["mask"]]
            bx, pr, ix = DO.yolo_decode(x, anchors)
            boxes.append(bx), probs.append(pr), idxs.append(ix)
            if capture is not None:
                capture[f"head{i}"] = x
        if i in keep:
            cache[i] = x
        if capture is not None:
            capture[i] = x
    bbox = torch.cat(boxes, dim=1)
    bbox[:, :, 2:4] = bbox[:, :, 2:4] / torch.tensor([net_info["width"], net_info["height"]])
    return {"bbox_xywh": bbox, "class_prob": torch.cat(probs, dim=1), "class_idx": torch.cat(idxs, dim=1)}
