"""Darknet model API — host-side mirror of ``yolov3/darknet.py`` of nrsyed/pytorch-yolov3.

Same public surface: ``parse_config``, ``blocks2modules``, ``DummyLayer``,
``MaxPool2d``, ``YOLOLayer`` and ``Darknet(config_fpath, device)`` with ``.blocks``,
``.net_info``, ``.modules_``, ``.device``, ``.blocks_to_cache``, ``.header``,
``.forward(x) -> dict`` and ``.load_weights(path) -> self``.  The ``torch.nn`` modules here only
HOLD parameters (so ``state_dict`` keys, ``.eval()``, ``.cuda()`` keep working); the arithmetic
runs in ``libyolov3_b200.so`` through a per-(batch, size) execution plan (``engine.py``): NHWC
bf16 activations, BatchNorm folded into the weights, shortcut / upsample / concat fused into
the tcgen05 convolution epilogues, the whole forward replayed as one CUDA graph.

There is no CPU path: ``forward`` raises unless it can run on an sm_100 CUDA device.
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .engine import Engine

# compiled plans kept per network: least-recently-used geometries (batch, height, width) beyond this
# many are dropped together with their buffers, graphs and staging memory (a video loop with a ragged
# last batch or a resolution sweep must not grow without bound)
MAX_GEOMETRIES = int(os.environ.get("Y3_MAX_GEOMETRIES", "6"))


class DummyLayer(torch.nn.Module):
    """Placeholder for route / shortcut blocks (reference: yolov3/darknet.py:6-13); the dataflow
    they describe is resolved by the execution plan."""


class MaxPool2d(torch.nn.MaxPool2d):
    """Max-pool with the reference's padding rule (yolov3/darknet.py:16-29): a stride-1 pool sees
    its input zero-padded on the right/bottom by ``kernel_size - 1``.  Standalone calls take NCHW
    float tensors like the reference module and run the CUDA kernel."""

    def forward(self, input_):
        """Deviation from the reference module (documented in INTEGRATION.md): the kernel works on the
        network's activation format, NHWC bf16 with channels in multiples of 8, so a stand-alone call
        rounds its input to bf16 (max of bf16 values is exact) and zero-pads the channel axis up to a
        multiple of 8 internally (the padding never reaches the result)."""
        dev = _lib.require_device(input_.device)
        n, c, h, w = input_.shape
        cp = (c + 7) // 8 * 8
        x = torch.zeros(n, h, w, cp, device=dev, dtype=torch.bfloat16)
        x[..., :c] = input_.to(dev).permute(0, 2, 3, 1)
        k, s = int(self.kernel_size), int(self.stride)
        ho, wo = (h, w) if (k > 1 and s == 1) else ((h - k) // s + 1, (w - k) // s + 1)
        y = torch.empty(n, ho, wo, cp, device=dev, dtype=torch.bfloat16)
        _lib.maxpool(x.data_ptr(), y.data_ptr(), n, h, w, cp, cp, cp, k, s)
        return y[..., :c].permute(0, 3, 1, 2).to(input_.dtype)


class YOLOLayer(torch.nn.Module):
    """YOLO head decode (reference: yolov3/darknet.py:32-122).  ``forward(x)`` takes the raw head
    tensor ``[B, A*(5+C), h, w]`` and returns ``(bbox_xywh [B,A*h*w,4], class_prob, class_idx)``
    with x,y in [0,1] and w,h in training-image pixels, exactly like the reference layer."""

    def __init__(self, anchors, mask, device="cuda"):
        super().__init__()
        self.anchors = [anchors[i] for i in mask]
        self.mask = mask
        self.device = device

    def forward(self, x):
        dev = _lib.require_device(x.device)
        b, p, h, w = x.shape
        a = len(self.anchors)
        classes = p // a - 5
        ld = (p + 3) // 4 * 4
        logits = torch.zeros(b, h, w, ld, device=dev, dtype=torch.float32)
        logits[..., :p] = x.to(dev, torch.float32).permute(0, 2, 3, 1)
        m = a * h * w
        bbox = torch.empty(b, m, 4, device=dev, dtype=torch.float32)
        prob = torch.empty(b, m, device=dev, dtype=torch.float32)
        idx = torch.empty(b, m, device=dev, dtype=torch.int64)
        # train size 1: the layer itself leaves w,h in training pixels (darknet.py:100-101)
        d = _lib.make_head_desc(b, h, w, self.anchors, classes, ld, 0, m, 1.0, 1.0)
        _lib.yolo_decode_dense(d, logits, bbox, prob, idx)
        return bbox, prob, idx


def _coerce(text):
    """``int`` if possible, else ``float``, else the raw string (darknet.py:163-176)."""
    try:
        return int(text)
    except ValueError:
        try:
            return float(text)
        except ValueError:
            return text


def parse_config(fpath):
    """Parse a Darknet ``.cfg`` into ``(blocks, net_info)`` (reference: yolov3/darknet.py:125-215).

    Reproduces the reference's conventions: lines that are blank or START with ``#`` are
    dropped before stripping; ``key=value`` has exactly one ``=``; comma-separated values become
    lists of coerced items; ``route.layers`` is always a list; ``anchors`` are grouped in
    ``[w, h]`` pairs; the ``[net]`` block is returned separately.
    """
    with open(fpath, "r") as f:
        lines = [ln.strip() for ln in f.readlines() if not (ln.isspace() or ln.startswith("#"))]
    blocks, net_info, block = [], None, None
    for ln in lines + ["["]:  # sentinel closes the last block
        if ln.startswith("["):
            if block is not None:
                if block["type"] == "net":
                    net_info = block
                else:
                    blocks.append(block)
            block = {"type": ln[1:-1]}
            continue
        key, raw = ln.split("=")
        key = key.strip()
        val = [_coerce(v.strip()) for v in raw.split(",")] if "," in raw else _coerce(raw.strip())
        if block["type"] == "route" and key == "layers" and isinstance(val, int):
            val = [val]
        if key == "anchors":
            val = [val[i:i + 2] for i in range(0, len(val), 2)]
        block[key] = val
    return blocks, net_info


def blocks2modules(blocks, net_info, device="cuda"):
    """``nn.ModuleList`` of ``nn.Sequential`` parameter holders, one per block, with the
    reference's submodule names ``conv_{i}``, ``batch_norm_{i}``, ``leaky_{i}``, ``maxpool_{i}``,
    ``route_{i}``, ``shortcut_{i}``, ``upsample_{i}``, ``yolo_{i}`` (yolov3/darknet.py:218-315)."""
    modules = torch.nn.ModuleList()
    prev_c = net_info["channels"]
    cur_c = None
    out_c = []
    for i, b in enumerate(blocks):
        seq = torch.nn.Sequential()
        t = b["type"]
        if t == "convolutional":
            bn = "batch_normalize" in b  # key presence, as in the reference (:237)
            k = b["size"]
            pad = (k - 1) // 2 if "pad" in b else 0  # key presence (:240)
            seq.add_module(f"conv_{i}", torch.nn.Conv2d(prev_c, b["filters"], k, stride=b["stride"], padding=pad,
                                                        bias=not bn))
            if bn:
                seq.add_module(f"batch_norm_{i}", torch.nn.BatchNorm2d(b["filters"]))
            if b["activation"] == "leaky":
                seq.add_module(f"leaky_{i}", torch.nn.LeakyReLU(0.1, inplace=True))
            # activation == "linear": identity (the reference builds a ReLU but never adds it, :258-261)
            cur_c = b["filters"]
        elif t == "maxpool":
            seq.add_module(f"maxpool_{i}", MaxPool2d(kernel_size=b["size"], stride=b["stride"]))
        elif t == "route":
            seq.add_module(f"route_{i}", DummyLayer())
            cur_c = sum(out_c[j] for j in b["layers"])  # negative j index from the end, like the reference
        elif t == "shortcut":
            seq.add_module(f"shortcut_{i}", DummyLayer())
            if b.get("activation") == "leaky":
                seq.add_module(f"leaky_{i}", torch.nn.LeakyReLU(0.1, inplace=True))
            assert cur_c == out_c[i + b["from"]], "shortcut operands must have equal channel counts"
        elif t == "upsample":
            seq.add_module(f"upsample_{i}", torch.nn.Upsample(scale_factor=b["stride"], mode="nearest"))
        elif t == "yolo":
            seq.add_module(f"yolo_{i}", YOLOLayer(b["anchors"], b["mask"], device=device))
        modules.append(seq)
        prev_c = cur_c
        out_c.append(cur_c)
    return modules


class Darknet(torch.nn.Module):
    """Darknet network built from a ``.cfg`` (reference: yolov3/darknet.py:318-476).

    Args:
        config_fpath (str): path to a Darknet .cfg (yolov3 / yolov3-tiny / yolov3-spp ...).
        device (str): CUDA device the network runs on, e.g. ``"cuda"`` or ``"cuda:1"``.
            ``"cpu"`` is accepted for construction / weight loading only.
    """

    def __init__(self, config_fpath, device="cuda"):
        super().__init__()
        self.blocks, self.net_info = parse_config(config_fpath)
        self.modules_ = blocks2modules(self.blocks, self.net_info, device=device)
        self.device = device
        self.header = None
        # outputs other blocks refer to; negative route indices become absolute IN PLACE (:334-349)
        self.blocks_to_cache = set()
        for i, b in enumerate(self.blocks):
            if b["type"] == "route":
                for j, idx in enumerate(b["layers"]):
                    if idx < 0:
                        b["layers"][j] = i + idx
                    self.blocks_to_cache.add(b["layers"][j])
            elif b["type"] == "shortcut":
                self.blocks_to_cache.update((i - 1, i + b["from"]))
        self._geometries = OrderedDict()  # (batch, height, width) -> {"engines": {slot: Engine}, ...}, LRU order
        self._weights_version = 0
        self._token = None

    # -- parameters ------------------------------------------------------------------------
    def load_weights(self, weights_path):
        """Read a Darknet ``.weights`` file into the parameter holders (reference:
        yolov3/darknet.py:407-476): 5 x int32 header (kept as ``self.header``) then float32
        values per convolutional block — ``[bn bias, bn weight, bn running_mean, bn running_var]``
        or ``[conv bias]``, then the OIHW kernel.  A short file raises ``RuntimeError``; trailing
        values are ignored, as in the reference.  Returns ``self``."""
        with open(weights_path, "rb") as f:
            self.header = np.fromfile(f, dtype=np.int32, count=5)
            flat = np.fromfile(f, dtype=np.float32)
        pos = 0

        def fill(dst):
            nonlocal pos
            n = dst.numel()
            if pos + n > flat.size:
                raise RuntimeError(f"{weights_path}: file ends after {flat.size} floats, block needs {pos + n}")
            dst.copy_(torch.from_numpy(flat[pos:pos + n]).view_as(dst))
            pos += n

        with torch.no_grad():
            for b, m in zip(self.blocks, self.modules_):
                if b["type"] != "convolutional":
                    continue
                conv = m[0]
                if "batch_normalize" in b and b["batch_normalize"]:
                    bn = m[1]
                    fill(bn.bias.data), fill(bn.weight.data), fill(bn.running_mean), fill(bn.running_var)
                else:
                    fill(conv.bias.data)
                fill(conv.weight.data)
        self.invalidate()
        return self

    def invalidate(self):
        """Drop compiled plans, folded weights and staging buffers.  Called by everything that can change
        the parameters behind the plans' back: ``load_weights``, ``load_state_dict``, ``.to()/.cuda()/
        .half()...`` (``_apply``), and by ``engine()`` itself when a parameter's version counter moved
        (in-place edits under ``torch.no_grad()``).  Writes through ``param.data`` bypass the version
        counter — call this by hand after such edits."""
        self.__dict__.setdefault("_geometries", OrderedDict()).clear()
        self.__dict__.pop("_folded_cache", None)
        self._weights_version = self.__dict__.get("_weights_version", 0) + 1
        self._token = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def _weights_token(self):
        """Cheap fingerprint of the parameters / BN statistics: sum of the tensors' version counters
        plus their identity (a replaced Parameter object changes it)."""
        tensors = self.__dict__.get("_tracked")
        if tensors is None or self._token is None:
            tensors = [t for t in list(self.parameters()) + list(self.buffers())]
            self.__dict__["_tracked"] = tensors
        return sum(t._version for t in tensors)

    def check_fresh(self):
        """Raise / refresh before running: plans must match the live parameters, and the network must
        be in eval mode — the plans fold BatchNorm's RUNNING statistics into the weights, which is what
        the reference computes after ``net.eval()`` (its CLI always calls it, yolov3/__main__.py:117);
        in train mode the reference would normalise with batch statistics instead."""
        if self.training and any(isinstance(m, torch.nn.BatchNorm2d) for m in self.modules()):
            raise RuntimeError("yolov3_b200.Darknet runs inference with BatchNorm folded into the weights: call "
                               "net.eval() first (train-mode batch statistics are not implemented)")
        tok = self._weights_token()
        if self._token is None:
            self._token = tok
        elif tok != self._token:
            self.invalidate()
            self._token = self._weights_token()

    # -- execution ------------------------------------------------------------------------------
    def _target_device(self):
        dev = torch.device(self.device)
        if dev.type != "cuda":
            p = next(self.parameters(), None)
            if p is not None and p.is_cuda:
                dev = p.device
        return _lib.require_device(dev)

    def engine(self, batch, height, width, slot=0, concurrent=False):
        """The compiled execution plan for this input geometry (built on first use).  ``slot`` > 0
        gives further independent instances (own buffers and graphs) of the same geometry:
        ``inference`` pipelines sub-batches through several of them.  ``concurrent`` (honoured when
        the plan is first built) marks a plan that runs next to others on different streams: its
        graphs are captured without programmatic dependent launch (see ``y3_set_pdl``).  Activation
        buffers are recycled along the network (liveness-based); set ``net.keep_activations = True``
        before the plan is built to give every block output its own buffer (debugging / tests that read
        intermediate tensors through ``engine.views``)."""
        geom = self.geometry(batch, height, width)
        eng = geom["engines"].get(slot)
        if eng is None:
            eng = Engine(self, batch, height, width, self._target_device(), pdl=not concurrent,
                         alias=not getattr(self, "keep_activations", False))
            geom["engines"][slot] = eng
        return eng

    def geometry(self, batch, height, width):
        """Per-geometry cache entry (plans by slot + whatever ``inference`` keeps beside them), most
        recently used last; the least recently used entries beyond ``MAX_GEOMETRIES`` are dropped."""
        key = (int(batch), int(height), int(width))
        geom = self._geometries.get(key)
        if geom is None:
            geom = self._geometries[key] = {"engines": {}}
            while len(self._geometries) > max(1, MAX_GEOMETRIES):
                self._geometries.popitem(last=False)
        else:
            self._geometries.move_to_end(key)
        return geom

    def forward(self, x):
        """Full forward + decode (reference: yolov3/darknet.py:351-405).

        Args:
            x: float tensor ``[B, 3, H, W]`` (RGB in [0,1]), H and W multiples of the network stride.
        Returns:
            dict with ``bbox_xywh`` float32 ``[B, M, 4]`` (cx, cy, w, h as fractions; w,h divided
            by the cfg's training size), ``class_prob`` float32 ``[B, M]``, ``class_idx`` int64 ``[B, M]``.
        """
        if x.dim() != 4 or x.shape[1] != self.net_info["channels"]:
            raise RuntimeError(f"expected input [B,{self.net_info['channels']},H,W], got {tuple(x.shape)}")
        self.check_fresh()
        eng = self.engine(x.shape[0], x.shape[2], x.shape[3])
        return eng.forward_dense(x)
