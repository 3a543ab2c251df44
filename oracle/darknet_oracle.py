"""CPU oracle for the model half of the hot path (TEST INFRASTRUCTURE — not product code).

A restatement of what nrsyed/pytorch-yolov3's ``yolov3/darknet.py`` computes, written
functionally (no nn.Module graph) on top of the same third-party arithmetic the reference
calls: ``torch`` CPU float32 ops (SURVEY.md §8c — the numerics live in PyTorch, which the
reference leaves unpinned; this image has torch 2.11.0).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package; the product (``pytorch-yolov3_b200/``) never does.

Pinned against the live reference by ``tests/golden/make_golden.py`` (run in the build
container, where /root/reference is importable) — see tests/test_oracle_golden.py.

Each function cites the reference lines it follows (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------
# cfg parsing — yolov3/darknet.py:125-215
# ---------------------------------------------------------------------------------------
def _coerce(text):
    """int, else float, else the string itself (darknet.py:163-176)."""
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            continue
    return text


def parse_config(path):
    """Darknet .cfg -> (blocks, net_info) with the reference's quirks (darknet.py:125-215):
    comment test before strip (:145-148), one '=' per line (:183), comma values become lists
    (:187-188), route.layers always a list (:195-200), anchors paired (:205-206), [net]
    returned separately (:210-213)."""
    with open(path, "r") as f:
        raw = f.readlines()
    lines = [ln.strip() for ln in raw if not (ln.isspace() or ln.startswith("#"))]
    blocks, net_info, cur = [], None, None

    def flush(b):
        nonlocal net_info
        if b is None:
            return
        if b["type"] == "net":
            net_info = b
        else:
            blocks.append(b)

    for ln in lines:
        if ln.startswith("["):
            flush(cur)
            cur = {"type": ln[1:-1]}
            continue
        key, val = ln.split("=")
        key = key.strip()
        if "," in val:
            val = [_coerce(v.strip()) for v in val.split(",")]
        else:
            val = _coerce(val.strip())
        if cur["type"] == "route" and key == "layers" and isinstance(val, int):
            val = [val]
        if key == "anchors":
            val = [val[i:i + 2] for i in range(0, len(val), 2)]
        cur[key] = val
    flush(cur)
    return blocks, net_info


def resolve_routes(blocks):
    """Absolute route indices, as Darknet.__init__ rewrites them in place (darknet.py:334-349).
    Returns the set of block outputs that must be kept."""
    keep = set()
    for i, b in enumerate(blocks):
        if b["type"] == "route":
            b["layers"] = [i + l if l < 0 else l for l in b["layers"]]
            keep.update(b["layers"])
        elif b["type"] == "shortcut":
            keep.add(i - 1)
            keep.add(i + b["from"])
    return keep


def conv_geometry(block):
    """(ksize, stride, pad, has_bn, leaky) of a [convolutional] block (darknet.py:236-261):
    pad=(k-1)//2 iff the key 'pad' is PRESENT; BN iff the key 'batch_normalize' is PRESENT;
    activation 'linear' is the identity (the ReLU object is created but never added)."""
    k = block["size"]
    pad = (k - 1) // 2 if "pad" in block else 0
    return k, block["stride"], pad, ("batch_normalize" in block), block["activation"] == "leaky"


def channel_plan(blocks, net_info):
    """Output channels of every block (darknet.py:224-313)."""
    out, prev = [], net_info["channels"]
    cur = None
    for i, b in enumerate(blocks):
        t = b["type"]
        if t == "convolutional":
            cur = b["filters"]
        elif t == "route":
            cur = sum(out[j if j >= 0 else i + j] for j in b["layers"])
        elif t == "shortcut":
            cur = out[i - 1]
        # maxpool / upsample / yolo keep the running channel count
        out.append(cur if cur is not None else prev)
        prev = out[-1]
    return out


# ---------------------------------------------------------------------------------------
# weights — yolov3/darknet.py:407-476
# ---------------------------------------------------------------------------------------
def conv_param_shapes(blocks, net_info):
    """[(block index, cin, cout, ksize, has_bn)] for every convolutional block, in file order."""
    chans = channel_plan(blocks, net_info)
    res, prev = [], net_info["channels"]
    for i, b in enumerate(blocks):
        if b["type"] == "convolutional":
            k, _, _, bn, _ = conv_geometry(b)
            res.append((i, prev, b["filters"], k, bn))
        prev = chans[i]
    return res


def read_weights(path, blocks, net_info):
    """Darknet .weights -> {block: dict(weight, bias | bn_*)} (darknet.py:416-475): 5 int32 header
    words, then per conv block [bn beta, bn gamma, bn mean, bn var] or [bias], then the OIHW
    kernel.  Trailing floats are ignored, as in the reference."""
    with open(path, "rb") as f:
        header = np.fromfile(f, dtype=np.int32, count=5)
        flat = np.fromfile(f, dtype=np.float32)
    params, p = {}, 0

    def take(n, shape):
        nonlocal p
        if p + n > flat.size:
            raise RuntimeError(f"weights file too short: need {p + n} floats, have {flat.size}")
        t = torch.from_numpy(flat[p:p + n].copy()).reshape(shape)
        p += n
        return t

    for i, cin, cout, k, bn in conv_param_shapes(blocks, net_info):
        d = {}
        if bn and blocks[i]["batch_normalize"]:
            d["bn_bias"] = take(cout, (cout,))
            d["bn_weight"] = take(cout, (cout,))
            d["bn_mean"] = take(cout, (cout,))
            d["bn_var"] = take(cout, (cout,))
        else:
            d["bias"] = take(cout, (cout,))
        d["weight"] = take(cout * cin * k * k, (cout, cin, k, k))
        params[i] = d
    return header, params


def write_weights(path, params, blocks, net_info, header=(0, 2, 0, 0, 0)):
    """Inverse of read_weights (the reference has no writer; tests need one)."""
    with open(path, "wb") as f:
        np.asarray(header, dtype=np.int32).tofile(f)
        for i, cin, cout, k, bn in conv_param_shapes(blocks, net_info):
            d = params[i]
            if "bn_bias" in d:
                for key in ("bn_bias", "bn_weight", "bn_mean", "bn_var"):
                    d[key].detach().cpu().numpy().astype(np.float32).tofile(f)
            else:
                d["bias"].detach().cpu().numpy().astype(np.float32).tofile(f)
            d["weight"].detach().cpu().numpy().astype(np.float32).tofile(f)


# ---------------------------------------------------------------------------------------
# forward — yolov3/darknet.py:351-405 and its modules
# ---------------------------------------------------------------------------------------
BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, darknet.py:252


def conv_block(x, block, prm):
    """Conv2d -> BatchNorm2d(eval) -> LeakyReLU(0.1) (darknet.py:244-257)."""
    k, stride, pad, bn, leaky = conv_geometry(block)
    y = F.conv2d(x, prm["weight"], prm.get("bias"), stride=stride, padding=pad)
    if "bn_bias" in prm:
        y = F.batch_norm(y, prm["bn_mean"], prm["bn_var"], prm["bn_weight"], prm["bn_bias"], False, 0.1, BN_EPS)
    if leaky:
        y = F.leaky_relu(y, 0.1)
    return y


def maxpool_block(x, block):
    """The reference's patched MaxPool2d (darknet.py:16-29): stride-1 pools first zero-pad the
    right and bottom by k-1; the pool itself has no padding and floor mode."""
    k, s = block["size"], block["stride"]
    if k > 1 and s == 1:
        x = F.pad(x, (0, k - 1, 0, k - 1), value=0.0)
    return F.max_pool2d(x, k, s)


def yolo_decode(x, anchors):
    """YOLOLayer.forward (darknet.py:48-122).  x [B, A*(5+C), h, w] float32; anchors = the masked
    (w, h) pairs.  Returns bbox_xywh [B, A*h*w, 4] (x,y in [0,1]; w,h in training pixels),
    class_prob [B, A*h*w], class_idx [B, A*h*w] int64.  Same torch ops in the same order."""
    A = len(anchors)
    B, P, h, w = x.shape
    C = P // A - 5
    x = x.reshape(B, A, C + 5, h, w)
    box = x[:, :, 0:4].clone()
    col = torch.linspace(0, w - 1, w).repeat(h, 1).to(x.device)  # the reference moves them with .to(device) (:83-84)
    row = torch.linspace(0, h - 1, h).repeat(w, 1).t().contiguous().to(x.device)
    box[:, :, 0].sigmoid_().add_(col).div_(w)
    box[:, :, 1].sigmoid_().add_(row).div_(h)
    anc = torch.tensor(anchors).to(x.device)  # ints in the cfg -> int64 tensor, as in the reference (:91-97)
    box[:, :, 2].exp_().mul_(anc[:, 0].reshape(1, A, 1, 1))
    box[:, :, 3].exp_().mul_(anc[:, 1].reshape(1, A, 1, 1))
    obj = x[:, :, 4:5].clone().sigmoid()
    cls = F.softmax(x[:, :, 5:].clone(), dim=2)
    prob, idx = torch.max(cls, 2, keepdim=True)
    prob = prob * obj
    return (box.permute(0, 1, 3, 4, 2).reshape(B, -1, 4), prob.reshape(B, -1), idx.reshape(B, -1))


def forward(x, blocks, net_info, params, capture=None):
    """Darknet.forward (darknet.py:351-405).  ``blocks`` must have absolute route indices
    (resolve_routes).  capture: optional dict filled with {block index: output tensor} for every
    block (and 'head<i>' -> raw logits fed to YOLO layer i) for teacher-forced comparisons."""
    keep = set()
    for i, b in enumerate(blocks):
        if b["type"] == "route":
            keep.update(b["layers"])
        elif b["type"] == "shortcut":
            keep.update((i - 1, i + b["from"]))
    cache, boxes, probs, idxs = {}, [], [], []
    for i, b in enumerate(blocks):
        t = b["type"]
        if t == "convolutional":
            x = conv_block(x, b, params[i])
        elif t == "maxpool":
            x = maxpool_block(x, b)
        elif t == "upsample":
            x = F.interpolate(x, scale_factor=b["stride"], mode="nearest")
        elif t == "route":
            x = torch.cat([cache[j] for j in b["layers"]], dim=1)
        elif t == "shortcut":
            x = cache[i - 1] + cache[i + b["from"]]
        elif t == "yolo":
            anchors = [b["anchors"][m] for m in b["mask"]]
            bx, pr, ix = yolo_decode(x, anchors)
            boxes.append(bx), probs.append(pr), idxs.append(ix)
            if capture is not None:
                capture[f"head{i}"] = x
        if i in keep:
            cache[i] = x
        if capture is not None:
            capture[i] = x
    bbox = torch.cat(boxes, dim=1)
    bbox[:, :, 2:4] = bbox[:, :, 2:4] / torch.tensor([net_info["width"], net_info["height"]]).to(bbox.device)
    return {"bbox_xywh": bbox, "class_prob": torch.cat(probs, dim=1), "class_idx": torch.cat(idxs, dim=1)}


# ---------------------------------------------------------------------------------------
# synthetic calibrated weights — SURVEY.md §8d (PyTorch-default init gives zero candidates)
# ---------------------------------------------------------------------------------------
def synth_params(blocks, net_info, size, seed=1234, calib_batch=2):
    """Seeded random weights whose activations neither vanish nor explode: He-style conv
    weights, BN gamma~U(0.8,1.2), beta~N(0,0.1), head bias~N(0,0.5); then one calibration
    forward sets every BN's running stats to its batch statistics and rescales each head conv
    so its pre-bias logits have std 2."""
    g = torch.Generator().manual_seed(seed)
    shapes = conv_param_shapes(blocks, net_info)
    params = {}
    for i, cin, cout, k, bn in shapes:
        fan_in = cin * k * k
        d = {"weight": torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (1.01 * fan_in))}
        if bn:
            d["bn_weight"] = torch.rand(cout, generator=g) * 0.4 + 0.8
            d["bn_bias"] = torch.randn(cout, generator=g) * 0.1
            d["bn_mean"] = torch.zeros(cout)
            d["bn_var"] = torch.ones(cout)
        else:
            d["bias"] = torch.randn(cout, generator=g) * 0.5
        params[i] = d
    x = torch.rand(calib_batch, net_info["channels"], size, size, generator=g)
    # calibration pass: same dataflow as forward(), statistics taken block by block
    keep = resolve_keep(blocks)
    cache = {}
    with torch.no_grad():
        for i, b in enumerate(blocks):
            t = b["type"]
            if t == "convolutional":
                prm = params[i]
                k, stride, pad, bn, leaky = conv_geometry(b)
                y = F.conv2d(x, prm["weight"], None, stride=stride, padding=pad)
                if bn:
                    prm["bn_mean"] = y.mean(dim=(0, 2, 3))
                    prm["bn_var"] = y.var(dim=(0, 2, 3), unbiased=False)
                else:
                    prm["weight"] = prm["weight"] * (2.0 / float(y.std()))
                x = conv_block(x, b, prm)
            elif t == "maxpool":
                x = maxpool_block(x, b)
            elif t == "upsample":
                x = F.interpolate(x, scale_factor=b["stride"], mode="nearest")
            elif t == "route":
                x = torch.cat([cache[j] for j in b["layers"]], dim=1)
            elif t == "shortcut":
                x = cache[i - 1] + cache[i + b["from"]]
            if i in keep:
                cache[i] = x
    return params


def resolve_keep(blocks):
    keep = set()
    for i, b in enumerate(blocks):
        if b["type"] == "route":
            keep.update(j if j >= 0 else i + j for j in b["layers"])
        elif b["type"] == "shortcut":
            keep.update((i - 1, i + b["from"]))
    return keep


def load_model(cfg_path):
    """(blocks with absolute routes, net_info)."""
    blocks, net_info = parse_config(cfg_path)
    resolve_routes(blocks)
    return blocks, net_info
