"""Execution plan for one (batch, height, width): the host-side "compiler" between the parsed
cfg and ``libyolov3_b200.so``.

What it decides once, at build time (reference semantics: yolov3/darknet.py:351-405):
  * NHWC bf16 buffers for every block output that is materialised; YOLO head convolutions
    write float32 logits.
  * BatchNorm folded into bf16 weights + fp32 bias (``W' = W*g/sqrt(var+eps)``,
    ``b' = beta - mean*g/sqrt(var+eps)``), repacked ``[Cout][R][S][Cin]`` for the implicit GEMM.
  * shortcut (:376-379) fused into the epilogue of the convolution that produces its first
    operand; nearest x2 upsample (:299-305) fused into the producing convolution's store;
    route / torch.cat (:369-375) made zero-copy by letting producers write into channel
    slices of a pre-allocated concat buffer; the three SPP max-pools run as one kernel.
    When a fusion's precondition fails (the intermediate has another consumer) the plan falls
    back to the stand-alone CUDA kernel for that block — still device code, never the CPU.
  * the kernel sequence is captured into CUDA graphs: (input packing + forward + dense decode)
    for ``Darknet.forward`` and (uint8 packing + forward + fused decode/threshold + NMS +
    compaction) for ``inference``.
"""
import os
import struct

import numpy as np
import torch

from . import _lib


def _p(buf):
    return buf.data_ptr() if buf.is_cuda else 0


def _ceil(x, m):
    return (x + m - 1) // m * m


class View:
    """NHWC tensor view: channel slice ``[c0, c0+C)`` of a buffer with pixel pitch ``ld``."""
    __slots__ = ("buf", "ptr", "C", "ld", "H", "W", "f32", "raw")

    def __init__(self, buf, ptr, C, ld, H, W, f32=False, raw=None):
        self.buf, self.ptr, self.C, self.ld, self.H, self.W, self.f32 = buf, ptr, C, ld, H, W, f32
        self.raw = raw  # the pooled allocation behind `buf` (identity = "same memory")


def _check_distinct(what, out, *ins):
    """Plan-build invariant of the activation pool: an op never writes the allocation it reads."""
    for v in ins:
        if v is not None and out.raw is not None and v.raw is out.raw and v.ptr == out.ptr:
            raise RuntimeError(f"internal error: {what} would write the buffer it reads (activation pool)")


class Engine:
    def __init__(self, net, batch, height, width, device, pdl=True, alias=True):
        self.net, self.B, self.H, self.W, self.device = net, batch, height, width, device
        # liveness-based reuse of activation memory (off: every block output keeps its own buffer, which
        # the teacher-forced tests need to read intermediate tensors back after a run)
        self.alias = alias and os.environ.get("Y3_NO_ALIAS", "0") != "1"
        # programmatic dependent launch in this plan's graphs (off for plans that run concurrently)
        self.pdl = pdl and os.environ.get("Y3_NO_PDL", "0") != "1"
        self.use_graphs = os.environ.get("Y3_NO_GRAPH", "0") != "1"
        self._graphs = {}
        self.conv_flops = 0  # algorithmic 2*MAC per batch, no padding credit
        self.conv_ops = []   # (block, launch closure, flops) for per-kernel timing in bench.py (uint8 programs)
        self.conv_ops_unfused_stem = []  # blocks 0-1 as separate launches (float32-input programs) when a stem exists
        self.conv_info = {}  # block -> GEMM shape and algorithmic HBM bytes of the launch (tools/conv_report.py)
        # device "meta" builds the plan without touching a GPU (host-logic tests); it cannot run
        self.dry = torch.device(device).type == "meta"
        if self.dry:
            self._build()
        else:
            with torch.cuda.device(device):
                self._build()

    # ------------------------------------------------------------------------------------
    # plan construction
    # ------------------------------------------------------------------------------------
    def _build(self):
        net, B, dev = self.net, self.B, self.device
        blocks = net.blocks
        nb = len(blocks)
        cin0 = net.net_info["channels"]

        # ---- shapes of every block output -------------------------------------------------
        shape = []  # (C, H, W)
        cur = (cin0, self.H, self.W)
        for i, b in enumerate(blocks):
            t = b["type"]
            if t == "convolutional":
                k, s = b["size"], b["stride"]
                pad = (k - 1) // 2 if "pad" in b else 0
                if k not in (1, 3) or s not in (1, 2):
                    raise NotImplementedError(f"block {i}: conv size={k} stride={s} not supported (1|3, 1|2)")
                cur = (b["filters"], (cur[1] + 2 * pad - k) // s + 1, (cur[2] + 2 * pad - k) // s + 1)
            elif t == "maxpool":
                k, s = b["size"], b["stride"]
                if not (k > 1 and s == 1):
                    cur = (cur[0], (cur[1] - k) // s + 1, (cur[2] - k) // s + 1)
            elif t == "upsample":
                if b["stride"] != 2:
                    raise NotImplementedError(f"block {i}: upsample stride {b['stride']} not supported (2)")
                cur = (cur[0], cur[1] * 2, cur[2] * 2)
            elif t == "route":
                srcs = [shape[j] for j in b["layers"]]
                if any(s_[1:] != srcs[0][1:] for s_ in srcs):
                    raise RuntimeError(f"block {i}: route sources have different spatial sizes")
                cur = (sum(s_[0] for s_ in srcs), srcs[0][1], srcs[0][2])
            elif t == "shortcut":
                if shape[i - 1] != shape[i + b["from"]]:
                    raise RuntimeError(f"block {i}: shortcut operands differ in shape")
                cur = shape[i - 1]
            elif t == "yolo":
                pass
            else:
                raise NotImplementedError(f"block {i}: type '{t}' not supported")
            shape.append(cur)

        # ---- aliases, producers, consumers --------------------------------------------------
        INPUT = -1
        alias = {}
        for i, b in enumerate(blocks):
            if b["type"] == "yolo":
                alias[i] = i - 1
            elif b["type"] == "route" and len(b["layers"]) == 1:
                alias[i] = b["layers"][0]

        def root(i):
            while i in alias:
                i = alias[i]
            return i

        def inputs_of(i):
            b = blocks[i]
            if b["type"] == "route":
                return [root(j) for j in b["layers"]]
            if b["type"] == "shortcut":
                return [root(i - 1), root(i + b["from"])]
            return [root(i - 1)] if i > 0 else [INPUT]

        consumers = {}
        for i in range(nb):
            if i in alias and blocks[i]["type"] != "yolo":
                continue
            for r in inputs_of(i):
                consumers.setdefault(r, []).append(i)

        def is_head(i):
            return blocks[i]["type"] == "convolutional" and i + 1 < nb and blocks[i + 1]["type"] == "yolo"

        # ---- fusion decisions ---------------------------------------------------------------------
        fused_into = {}   # conv block -> shortcut / upsample block whose output it writes
        residual_of = {}  # conv block -> root block providing the residual
        for i, b in enumerate(blocks):
            if b["type"] == "shortcut":
                a, r = root(i - 1), root(i + b["from"])
                if (a >= 0 and blocks[a]["type"] == "convolutional" and not is_head(a) and consumers.get(a) == [i]
                        and a != r and a not in fused_into):
                    fused_into[a] = i
                    residual_of[a] = r
            elif b["type"] == "upsample":
                a = root(i - 1)
                if (a >= 0 and blocks[a]["type"] == "convolutional" and not is_head(a) and consumers.get(a) == [i]
                        and a not in fused_into):
                    fused_into[a] = i
        fused_blocks = set(fused_into.values())

        # ---- concat placement ------------------------------------------------------------------------
        placement = {}  # root block -> (route block, channel offset)
        concat_buf = {}
        for i, b in enumerate(blocks):
            if b["type"] == "route" and len(b["layers"]) > 1:
                C, H, W = shape[i]
                if C % 8:
                    raise NotImplementedError(f"block {i}: concat of {C} channels (multiple of 8 required)")
                concat_buf[i] = None  # allocated from the pool when its first producer runs (concat_of)
                off = 0
                for j in b["layers"]:
                    r = root(j)
                    movable = (r >= 0 and r not in placement and r not in concat_buf and not is_head(r)
                               and off % 8 == 0)
                    if movable:
                        placement[r] = (i, off)
                    off += shape[j][0]

        views = {}
        self._keep = []  # buffers kept alive

        # ---- activation memory: a pool of buffers reused once their last reader has been emitted --------
        # Kernels run in emission order on one stream (every kernel waits for its predecessors before it
        # touches activations, also under programmatic dependent launch), so a buffer whose last consumer
        # sits at an EARLIER block index than the producer being emitted can be handed out again.  Best fit
        # over the free buffers, a new one when none is large enough: Darknet's tensors shrink with depth,
        # so the few large early buffers serve the whole network (5.2 GB -> ~1 GB at batch 64, 416x416).
        FOREVER = 1 << 30
        pool = []  # [uint8 tensor, free?, last consumer block]
        self.activation_bytes_unaliased = 0

        def last_reader(j):
            return max(consumers.get(j, [FOREVER]))

        def pool_get(nbytes, last):
            nbytes = _ceil(nbytes, 1024)
            self.activation_bytes_unaliased += nbytes
            best = None
            if self.alias and last < FOREVER:
                for e in pool:
                    if e[1] and e[0].numel() >= nbytes and (best is None or e[0].numel() < best[0].numel()):
                        best = e
            if best is None:
                best = [torch.empty(nbytes, device=dev, dtype=torch.uint8), False, last]
                pool.append(best)
            best[1], best[2] = False, last
            return best[0]

        def pool_release(i):
            if self.alias:
                for e in pool:
                    if not e[1] and e[2] < i:
                        e[1] = True

        def typed(raw, dims, dtype):
            n = 1
            for d_ in dims:
                n *= d_
            return raw[:n * (4 if dtype == torch.float32 else 2)].view(dtype).view(*dims)

        concat_raw = {}

        def concat_of(r):
            if concat_buf[r] is None:
                C, H, W = shape[r]
                concat_raw[r] = pool_get(B * H * W * C * 2, last_reader(r))
                concat_buf[r] = typed(concat_raw[r], (B, H, W, C), torch.bfloat16)
            return concat_buf[r]

        def alloc(i, f32=False, c_store=None):
            C, H, W = shape[i]
            if i in placement and not f32:
                r, off = placement[i]
                buf = concat_of(r)
                return View(buf, _p(buf) + off * 2, C, buf.shape[3], H, W, raw=concat_raw[r])
            cs = c_store or C
            dtype = torch.float32 if f32 else torch.bfloat16
            # YOLO head logits are read by the decode kernels AFTER the whole backbone: never recycled
            last = FOREVER if f32 else last_reader(i)
            raw = pool_get(B * H * W * cs * (4 if f32 else 2), last)
            buf = typed(raw, (B, H, W, cs), dtype)
            return View(buf, _p(buf), cs, cs, H, W, f32, raw=raw)

        # network input.  A first layer that is 3x3/s1/pad1 over <= 3 channels (every shipped cfg)
        # gets its im2col done by the packing kernel: the input buffer holds K = 27 -> 32 taps
        # per pixel and block 0 runs as a plain K=32 GEMM.  Otherwise channels are padded to 16.
        b0 = blocks[0]
        self.first_im2col = (b0["type"] == "convolutional" and b0["size"] == 3 and b0["stride"] == 1
                             and "pad" in b0 and 9 * cin0 <= 32
                             and os.environ.get("Y3_NO_FIRST_IM2COL", "0") != "1")
        cin_pad = 32 if self.first_im2col else _ceil(cin0, 16)
        in_buf = torch.zeros(B, self.H, self.W, cin_pad, device=dev, dtype=torch.bfloat16)
        views[INPUT] = View(in_buf, _p(in_buf), cin_pad, cin_pad, self.H, self.W)
        self.in_view = views[INPUT]
        # channels of the input buffer that hold the image itself (centre tap when im2col'ed)
        self.input_image_channels = (4 * cin0, 5 * cin0) if self.first_im2col else (0, cin0)

        ops = []          # (closure, group) executed in order; group None = always,
        self.op_names = []  # "stem_unfused" / "stem_fused" = alternative forms of blocks 0-1 (see run_backbone)
        self.op_groups = []

        self.mem_ops = []  # (name, launch closure, algorithmic HBM bytes) of the memory-bound block kernels

        def emit(name, fn, group=None, hbm_bytes=None):
            ops.append(fn)
            self.op_names.append(name)
            self.op_groups.append(group)
            if hbm_bytes is not None:
                self.mem_ops.append((name, fn, hbm_bytes))

        use_chain = os.environ.get("Y3_NO_CHAIN", "0") != "1"
        use_fused_decode = os.environ.get("Y3_NO_FUSED_DECODE", "0") != "1"
        self.num_fused_heads = 0

        def conv_geom(j):
            bj = blocks[j]
            return bj["size"], bj["stride"], ((bj["size"] - 1) // 2 if "pad" in bj else 0)

        def res_chain_at(i):
            """conv1x1(64->32) -> conv3x3/1(32->64) -> shortcut back to the 1x1's input: one kernel."""
            if not use_chain or i + 2 >= nb or blocks[i + 1]["type"] != "convolutional":
                return False
            t = fused_into.get(i + 1)
            if t is None or blocks[t]["type"] != "shortcut" or consumers.get(i) != [i + 1] or i in fused_into:
                return False
            src = inputs_of(i)[0]
            if src == INPUT or residual_of.get(i + 1) != src or views[src].C != 64 or views[src].f32:
                return False
            if conv_geom(i) != (1, 1, 0) or conv_geom(i + 1) != (3, 1, 1):
                return False
            if blocks[i]["filters"] != 32 or blocks[i + 1]["filters"] != 64 or is_head(i) or is_head(i + 1):
                return False
            return views[src].H % 16 == 0 and views[src].W % 8 == 0

        # blocks 0-1 as conv3x3(3->32) -> conv3x3/2(32->64) straight from uint8 images: one kernel
        self.stem = None
        stem_ok = (use_chain and self.first_im2col and cin0 == 3 and nb > 2 and b0["filters"] == 32
                   and blocks[1]["type"] == "convolutional" and conv_geom(1) == (3, 2, 1)
                   and blocks[1]["filters"] == 64 and consumers.get(0) == [1] and 0 not in fused_into
                   and 1 not in fused_into and not is_head(1) and self.H % 32 == 0 and self.W % 32 == 0)

        folded = self._folded_weights

        done = set()
        heads = []
        for i, b in enumerate(blocks):
            t = b["type"]
            if i in done or i in fused_blocks:
                continue
            pool_release(i)
            if t == "convolutional" and res_chain_at(i):
                xin = views[inputs_of(i)[0]]
                tgt = fused_into[i + 1]
                yv = alloc(tgt)
                views[tgt] = views[i + 1] = yv
                _check_distinct(f"chain{i}", yv, xin)
                w1, b1 = folded(i, 64, 32)
                w2, b2 = folded(i + 1, 32, 64)
                fn = (lambda xp=xin.ptr, w1=w1, b1=b1, w2=w2, b2=b2, yp=yv.ptr, h=xin.H, w=xin.W, lx=xin.ld, ly=yv.ld,
                      l1=blocks[i]["activation"] == "leaky", l2=blocks[i + 1]["activation"] == "leaky":
                      _lib.conv_chain_res64(xp, w1, b1, w2, b2, yp, n=B, h=h, w=w, ld_x=lx, ld_y=ly, leaky1=l1, leaky2=l2))
                emit(f"chain{i}", fn)
                flops = 2 * B * xin.H * xin.W * (32 * 64 + 64 * 32 * 9)
                self.conv_flops += flops
                self.conv_ops.append((i, fn, flops))
                px = B * xin.H * xin.W
                self.conv_info[i] = {"kind": "res chain 1x1+3x3+add", "M": px, "N": 64, "K": 64 + 9 * 32, "k": 3, "s": 1,
                                     "bytes": px * 64 * 2 * 2 + (64 * 32 + 9 * 32 * 64) * 2}
                done.add(i + 1)
            elif t == "convolutional":
                xin = views[inputs_of(i)[0]]
                k, s = b["size"], b["stride"]
                pad = (k - 1) // 2 if "pad" in b else 0
                head = is_head(i)
                cout = b["filters"]
                if head:
                    if consumers.get(i, []) != [i + 1]:
                        raise NotImplementedError(f"block {i}: YOLO head conv with extra consumers")
                    cout_pad = _ceil(cout, 16)
                else:
                    if cout % 16:
                        raise NotImplementedError(f"block {i}: filters={cout} must be a multiple of 16")
                    cout_pad = cout
                if xin.C % 16:
                    raise NotImplementedError(f"block {i}: input channels {xin.C} must be a multiple of 16")
                tgt = fused_into.get(i, i)
                up = blocks[tgt]["type"] == "upsample"
                yv = alloc(tgt, f32=head, c_store=cout_pad if head else None)
                views[tgt] = yv
                if tgt != i:
                    views[i] = yv  # never read (single consumer), kept for introspection
                res = views[residual_of[i]] if i in residual_of else None
                _check_distinct(f"conv{i}", yv, xin, res)
                as_gemm = self.first_im2col and i == 0
                w, bias = folded(i, xin.C, cout_pad, flatten_taps=as_gemm)
                kk, pp = (1, 0) if as_gemm else (k, pad)
                kw = dict(n=B, h=xin.H, w_in=xin.W, cin=xin.C, cout=cout_pad, ksize=kk, stride=s, pad=pp,
                          ld_x=xin.ld, ld_y=yv.ld, leaky=b["activation"] == "leaky",
                          res_ptr=res.ptr if res else None, ld_res=res.ld if res else 0, out_f32=head,
                          upsample2x=up)
                fn = (lambda xp=xin.ptr, w=w, bias=bias, yp=yv.ptr, kw=kw: _lib.conv2d(xp, w, bias, yp, **kw))
                in_stem = stem_ok and i in (0, 1)
                # YOLO head with 3 anchors x 80 classes: the detection programs decode in the conv epilogue
                yb = blocks[i + 1] if head else None
                fuse_head = (head and use_fused_decode and k == 1 and s == 1 and xin.C % 64 == 0 and cout == 255
                             and len(yb["mask"]) == 3 and yb["classes"] == 80)
                emit(f"conv{i}", fn, "stem_unfused" if in_stem else ("head_logits" if fuse_head else None))
                if fuse_head:
                    hfn = (lambda xp=xin.ptr, w=w, bias=bias, hk=len(heads), h=xin.H, wi=xin.W, c=xin.C, lx=xin.ld:
                           _lib.conv2d_yolo_head(xp, w, bias, self.head_descs[hk][0], 0.0, self.orig_hw,
                                                 self.cands, self.counts, self.cap, n=B, h=h, w_in=wi, cin=c, ld_x=lx,
                                                 dev_thresholds=self.thresh))
                    emit(f"headconv{i}", hfn, "head_fused")
                    self.num_fused_heads += 1
                ho, wo = shape[i][1], shape[i][2]
                cin_real = cin0 if inputs_of(i)[0] == INPUT else shape[inputs_of(i)[0]][0]
                flops = 2 * B * ho * wo * cout * cin_real * k * k
                self.conv_flops += flops
                self.conv_info[i] = {"kind": "conv", "M": B * ho * wo, "N": cout, "K": cin_real * k * k, "k": k, "s": s,
                                     "bytes": (B * xin.H * xin.W * xin.C * 2 + B * ho * wo * cout * (4 if head else 2)
                                               * (4 if up else 1) + cout * cin_real * k * k * 2
                                               + (B * ho * wo * cout * 2 if res else 0))}
                if in_stem:
                    self.conv_ops_unfused_stem.append((i, fn, flops))
                else:
                    self.conv_ops.append((i, hfn if fuse_head else fn, flops))
                if in_stem and i == 1:
                    (w1, b1), (w2, b2) = self._stem_w0, (w, bias)
                    sfn = (lambda w1=w1, b1=b1, w2=w2, b2=b2, yp=yv.ptr, ly=yv.ld,
                           l1=blocks[0]["activation"] == "leaky", l2=b["activation"] == "leaky":
                           _lib.conv_chain_stem_u8(self.in_u8, w1, b1, w2, b2, yp, ld_y=ly, leaky1=l1, leaky2=l2))
                    emit("stem0", sfn, "stem_fused")
                    self.stem = sfn
                    self.conv_ops.insert(0, (0, sfn, sum(f for _, _, f in self.conv_ops_unfused_stem)))
                    self.conv_info["stem"] = {"kind": "uint8 stem 3x3+3x3/2", "M": B * ho * wo, "N": 64, "K": 27 * 4 + 288,
                                              "k": 3, "s": 2, "bytes": B * self.H * self.W * 3 + B * ho * wo * 64 * 2}
                elif in_stem:
                    self._stem_w0 = self._stem_first_weights(bias)
                if head:
                    heads.append((i + 1, yv))
            elif t == "maxpool":
                src = inputs_of(i)[0]
                xin = views[src]
                k, s = b["size"], b["stride"]
                if xin.C % 8:
                    raise NotImplementedError(f"block {i}: maxpool over {xin.C} channels (multiple of 8 required)")
                # SPP: k5/k9/k13 stride-1 pools of the same tensor -> one kernel
                trio = None
                if s == 1 and k == 5:
                    sib = {blocks[j]["size"]: j for j in range(i + 1, nb)
                           if blocks[j]["type"] == "maxpool" and blocks[j]["stride"] == 1
                           and inputs_of(j)[0] == src and blocks[j]["size"] in (9, 13)}
                    if set(sib) == {9, 13}:
                        trio = (i, sib[9], sib[13])
                if trio:
                    vs = [alloc(j) for j in trio]
                    for j, v in zip(trio, vs):
                        views[j] = v
                        done.add(j)
                    if len({v.ld for v in vs}) == 1:
                        emit(f"spp{i}", lambda xp=xin.ptr, a=vs[0].ptr, b_=vs[1].ptr, c=vs[2].ptr, h=xin.H, w=xin.W,
                             C=xin.C, lx=xin.ld, ly=vs[0].ld: _lib.spp3(xp, a, b_, c, B, h, w, C, lx, ly),
                             hbm_bytes=B * xin.H * xin.W * xin.C * 2 * 4)  # one read, three pooled writes
                    else:
                        for j, v in zip(trio, vs):
                            emit(f"maxpool{j}", lambda xp=xin.ptr, yp=v.ptr, h=xin.H, w=xin.W, C=xin.C, lx=xin.ld,
                                 ly=v.ld, kk=blocks[j]["size"]: _lib.maxpool(xp, yp, B, h, w, C, lx, ly, kk, 1),
                                 hbm_bytes=B * xin.H * xin.W * xin.C * 2 * 2)
                else:
                    yv = alloc(i)
                    views[i] = yv
                    emit(f"maxpool{i}", lambda xp=xin.ptr, yp=yv.ptr, h=xin.H, w=xin.W, C=xin.C, lx=xin.ld, ly=yv.ld,
                         kk=k, ss=s: _lib.maxpool(xp, yp, B, h, w, C, lx, ly, kk, ss),
                         hbm_bytes=B * xin.C * 2 * (xin.H * xin.W + yv.H * yv.W))
            elif t == "upsample":  # not fused: stand-alone kernel
                xin = views[inputs_of(i)[0]]
                yv = alloc(i)
                views[i] = yv
                emit(f"upsample{i}", lambda xp=xin.ptr, yp=yv.ptr, h=xin.H, w=xin.W, C=xin.C, lx=xin.ld, ly=yv.ld:
                     _lib.upsample2x(xp, yp, B, h, w, C, lx, ly), hbm_bytes=B * xin.H * xin.W * xin.C * 2 * 5)
            elif t == "shortcut":  # not fused: stand-alone add
                a, r = (views[j] for j in inputs_of(i))
                yv = alloc(i)
                views[i] = yv
                emit(f"add{i}", lambda ap=a.ptr, bp=r.ptr, yp=yv.ptr, px=B * a.H * a.W, C=a.C, la=a.ld, lb=r.ld,
                     ly=yv.ld: _lib.add(ap, bp, yp, px, C, la, lb, ly), hbm_bytes=B * a.H * a.W * a.C * 2 * 3)
            elif t == "route":
                if len(b["layers"]) == 1:
                    views[i] = views[root(i)]
                    continue
                buf = concat_of(i)
                off = 0
                for j in b["layers"]:
                    r = root(j)
                    if placement.get(r) != (i, off):  # source lives elsewhere: copy its channels in
                        sv = views[r]
                        if sv.f32 or sv.C % 8 or off % 8:
                            raise NotImplementedError(f"block {i}: cannot concatenate source block {j}")
                        emit(f"copy{i}_{j}", lambda xp=sv.ptr, yp=_p(buf) + off * 2, px=B * sv.H * sv.W,
                             C=shape[j][0], lx=sv.ld, ly=buf.shape[3]: _lib.copy_channels(xp, yp, px, C, lx, ly),
                             hbm_bytes=B * sv.H * sv.W * shape[j][0] * 2 * 2)
                    off += shape[j][0]
                views[i] = View(buf, _p(buf), shape[i][0], buf.shape[3], shape[i][1], shape[i][2])
            elif t == "yolo":
                views[i] = views[root(i)]
        self.views = views
        self._keep = [e[0] for e in pool]
        self.activation_bytes = sum(e[0].numel() for e in pool)
        self.backbone_ops = ops
        self.num_fused = {"shortcut": len(residual_of), "upsample": len(fused_into) - len(residual_of),
                          "concat_slices": len(placement)}

        # ---- decode -----------------------------------------------------------------------------------
        if not heads:
            raise RuntimeError("cfg has no [yolo] block")
        self.head_descs = []
        fused_names = {n for n, g_ in zip(self.op_names, self.op_groups) if g_ == "head_fused"}
        self.head_fused = [f"headconv{y - 1}" in fused_names for y, _ in heads]
        off = 0
        M = sum(len(blocks[y]["mask"]) * v.H * v.W for y, v in heads)
        classes = None
        for y, v in heads:
            yb = blocks[y]
            anchors = [yb["anchors"][m] for m in yb["mask"]]
            nc = shape[y - 1][0] // len(anchors) - 5
            classes = nc if classes is None else classes
            if nc != classes:
                raise NotImplementedError("YOLO heads with different class counts")
            d = _lib.make_head_desc(B, v.H, v.W, anchors, nc, v.ld, off, M, net.net_info["width"],
                                    net.net_info["height"])
            self.head_descs.append((d, v.buf))
            off += len(anchors) * v.H * v.W
        self.M, self.num_classes = M, classes

        # static I/O buffers
        self.in_f32 = torch.zeros(B, cin0, self.H, self.W, device=dev, dtype=torch.float32)
        self.in_u8 = torch.zeros(B, self.H, self.W, 3, device=dev, dtype=torch.uint8) if cin0 == 3 else None
        self.bbox = torch.empty(B, M, 4, device=dev, dtype=torch.float32)
        self.prob = torch.empty(B, M, device=dev, dtype=torch.float32)
        self.cidx = torch.empty(B, M, device=dev, dtype=torch.int64)
        self.cap = M
        self.orig_hw = torch.zeros(B, 2, device=dev, dtype=torch.int32)
        self.cands = torch.zeros(B, M, 8, device=dev, dtype=torch.int32)
        self.counts = torch.zeros(B, device=dev, dtype=torch.int32)
        self.sorted = torch.zeros(B, M, 8, device=dev, dtype=torch.int32)
        self.keep = torch.zeros(B, M, device=dev, dtype=torch.uint8)
        self.dets = torch.zeros(B * M, 8, device=dev, dtype=torch.int32)
        # everything the host needs to know about a finished batch in ONE int32 block (one D2H copy):
        # [detections kept per image (B) | their total (1) | kept per (image, class) | first box per (image, class)]
        self.meta = torch.zeros(B + 1 + 2 * B * classes, device=dev, dtype=torch.int32)
        self.det_counts = self.meta[:B]
        self.det_counts_total = self.meta[:B + 1]
        self.seg_meta = self.meta[B + 1:].view(2, B, classes)
        self.class_kept, self.first_box = self.seg_meta[0], self.seg_meta[1]
        # y3_thresholds record (prob_thresh f32, pad, iou_thresh f64): the kernels read the thresholds from
        # here at run time, so one captured graph per program serves every threshold setting
        self.thresh = torch.zeros(16, device=dev, dtype=torch.uint8)
        self._thresh_host = None
        self.class_start = torch.zeros(B, classes + 1, device=dev, dtype=torch.int32)
        self.dst_off = torch.zeros(B, classes, device=dev, dtype=torch.int32)
        self.out_tlbr = torch.empty(B * M, 4, device=dev, dtype=torch.int64)
        self.out_prob = torch.empty(B * M, device=dev, dtype=torch.float32)
        self.out_cls = torch.empty(B * M, device=dev, dtype=torch.int64)
        ws_bytes = 0 if self.dry else _lib.nms_workspace_bytes(B, M, classes)
        self.nms_ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)

    def _folded_weights(self, i, cin_store, cout_store, flatten_taps=False):
        """BN-folded bf16 ``[cout_store][R][S][cin_store]`` weights + fp32 bias of conv block ``i``
        (shared between plans of one network; a4 in SURVEY.md §8a).  flatten_taps: the (r, s, c)
        taps become ONE K axis padded to cin_store (first layer run as a GEMM on im2col'ed input)."""
        cache = self.net.__dict__.setdefault("_folded_cache", {})
        key = (self.net._weights_version, i, cin_store, cout_store, str(self.device), flatten_taps)
        if key in cache:
            return cache[key]
        seq = self.net.modules_[i]
        conv = seq[0]
        with torch.no_grad():
            W = conv.weight.detach().to(self.device, torch.float32)
            if len(seq) > 1 and isinstance(seq[1], torch.nn.BatchNorm2d):
                bn = seq[1]
                scale = bn.weight.detach().to(self.device, torch.float32) / torch.sqrt(
                    bn.running_var.detach().to(self.device, torch.float32) + bn.eps)
                W = W * scale.view(-1, 1, 1, 1)
                bias = bn.bias.detach().to(self.device, torch.float32) - \
                    bn.running_mean.detach().to(self.device, torch.float32) * scale
            else:
                bias = conv.bias.detach().to(self.device, torch.float32)
            cout, cin, k, _ = W.shape
            if flatten_taps:
                Wk = torch.zeros(cout_store, 1, 1, cin_store, device=self.device, dtype=torch.float32)
                Wk[:cout, 0, 0, :k * k * cin] = W.permute(0, 2, 3, 1).reshape(cout, k * k * cin)
            else:
                Wk = torch.zeros(cout_store, k, k, cin_store, device=self.device, dtype=torch.float32)
                Wk[:cout, :, :, :cin] = W.permute(0, 2, 3, 1)
            bf = torch.zeros(cout_store, device=self.device, dtype=torch.float32)
            bf[:cout] = bias
            out = (Wk.to(torch.bfloat16).contiguous(), bf.contiguous())
        # drop entries of older weight versions
        for k_ in [k_ for k_ in cache if k_[0] != self.net._weights_version]:
            del cache[k_]
        cache[key] = out
        return out

    def _stem_first_weights(self, bias):
        """Block 0 (3x3 over RGB) for the fused uint8 stem: bf16 ``[32][3 filter rows][16]`` with the
        9 taps of a filter row in (column, BGR byte) order — the image's byte order — padded to 16."""
        cache = self.net.__dict__.setdefault("_folded_cache", {})
        key = (self.net._weights_version, "stem0", str(self.device))
        if key not in cache:
            seq = self.net.modules_[0]
            with torch.no_grad():
                W = seq[0].weight.detach().to(self.device, torch.float32)  # [32, rgb, dy, dx]
                if len(seq) > 1 and isinstance(seq[1], torch.nn.BatchNorm2d):
                    bn = seq[1]
                    W = W * (bn.weight.detach().to(self.device, torch.float32) / torch.sqrt(
                        bn.running_var.detach().to(self.device, torch.float32) + bn.eps)).view(-1, 1, 1, 1)
                Wk = torch.zeros(W.shape[0], 3, 16, device=self.device, dtype=torch.float32)
                # [n, dy, dx, byte] with byte b = BGR position -> channel 2 - b
                Wk[:, :, :9] = W.flip(1).permute(0, 2, 3, 1).reshape(W.shape[0], 3, 9)
                cache[key] = Wk.to(torch.bfloat16).contiguous()
        return cache[key], bias

    # ------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------
    def run_backbone(self, fused_stem=False, fused_heads=False):
        """All block launches in order.  Blocks 0-1 exist in two forms when the plan has a stem:
        separate convolutions over the packed input buffer (float32 input), or the fused
        uint8-image kernel (`fused_stem`, uint8 programs — no packing launch needed).  YOLO head
        convolutions exist in two forms as well: float32 logits for `Darknet.forward`, or
        (`fused_heads`, detection programs) decode + threshold + candidate append in the epilogue —
        the caller zeroes `counts` and sets `_prob_thresh` first."""
        skip = {"stem_unfused" if (fused_stem and self.stem is not None) else "stem_fused",
                "head_logits" if fused_heads else "head_fused"}
        for op, group in zip(self.backbone_ops, self.op_groups):
            if group not in skip:
                op()

    def _decode_dense(self):
        for d, logits in self.head_descs:
            _lib.yolo_decode_dense(d, logits, self.bbox, self.prob, self.cidx)

    def set_thresholds(self, prob_thresh, iou_thresh):
        """Update the device-resident ``y3_thresholds`` record (stream-ordered; a no-op when unchanged).
        The source is pageable memory on purpose: the copy is staged before the call returns, so a later
        update cannot overtake an earlier one that is still queued."""
        want = (float(prob_thresh), float(iou_thresh))
        if want != self._thresh_host:
            rec = np.frombuffer(struct.pack("<ffd", want[0], 0.0, want[1]), dtype=np.uint8).copy()
            self.thresh.copy_(torch.from_numpy(rec), non_blocking=True)
            self._thresh_host = want

    def _detect(self, prob_thresh=None, iou_thresh=None, compact=True, fused_stem=False, emit=False):
        """Backbone + decode + NMS (+ compaction / final arrays).  Heads that can decode in their epilogue
        do; the others (other class / anchor counts) write logits and run the stand-alone decode kernel.
        Thresholds come from the device record; passing them here (eager callers only, never inside a
        graph capture) updates it first."""
        if prob_thresh is not None:
            self.set_thresholds(prob_thresh, iou_thresh)
        self.counts.zero_()
        self.run_backbone(fused_stem=fused_stem, fused_heads=True)
        for hk, (d, logits) in enumerate(self.head_descs):
            if not self.head_fused[hk]:
                _lib.yolo_decode_cands(d, logits, 0.0, self.orig_hw, self.cands, self.counts, self.cap,
                                       dev_thresholds=self.thresh)
        self._nms_tail(compact=compact, emit=emit)

    def _detect_tail(self, prob_thresh, iou_thresh, compact=True):
        """Decode (stand-alone kernels, from the logits of a previous run_backbone()) + NMS."""
        self.set_thresholds(prob_thresh, iou_thresh)
        self.counts.zero_()
        for d, logits in self.head_descs:
            _lib.yolo_decode_cands(d, logits, 0.0, self.orig_hw, self.cands, self.counts, self.cap,
                                   dev_thresholds=self.thresh)
        self._nms_tail(compact=compact)

    def _nms_tail(self, compact=True, emit=False):
        _lib.nms(self.cands, self.counts, self.B, self.cap, self.num_classes, 0.0, 1, self.sorted, self.keep,
                 self.first_box, self.nms_ws, class_start=self.class_start, class_kept=self.class_kept,
                 dev_thresholds=self.thresh)
        if compact:  # flat y3_cand records (device-resident result; bench / multi-GPU gather)
            _lib.compact_kept(self.sorted, self.keep, self.counts, self.B, self.cap, self.dets, self.det_counts, 1)
        if emit:  # the reference's final arrays, class groups ascending inside an image (see inference_batches)
            _lib.plan_destinations(self.class_kept, self.B, self.num_classes, self.dst_off, self.det_counts_total)
            self.emit()

    def emit(self, out=None):
        """Second, tiny launch of `inference`: write the kept records as the reference's final arrays
        at the per-class-group positions the host put into ``dst_off`` — into this plan's own
        output buffers, or into ``out = (tlbr, prob, cls)`` shared by the sub-batches of one call."""
        tlbr, prob, cls = out if out is not None else (self.out_tlbr, self.out_prob, self.out_cls)
        _lib.emit_detections(self.sorted, self.keep, self.class_start, self.dst_off, self.B, self.cap,
                             self.num_classes, tlbr, prob, cls)

    def _program(self, key):
        kind = key[0]
        pack_f32 = _lib.im2col3x3_nchw_f32 if self.first_im2col else _lib.pack_nchw_f32
        pack_u8 = _lib.im2col3x3_bgr_u8 if self.first_im2col else _lib.pack_bgr_u8
        if kind == "dense_f32":
            def fn():
                pack_f32(self.in_f32, self.in_view.buf, self.in_view.C)
                self.run_backbone()
                self._decode_dense()
        elif kind == "det_u8":
            def fn():
                if self.stem is None:
                    pack_u8(self.in_u8, self.in_view.buf, self.in_view.C)
                self._detect(fused_stem=True)
        elif kind == "nms_u8":  # inference(): final arrays are emitted by a second launch (Engine.emit)
            def fn():
                if self.stem is None:
                    pack_u8(self.in_u8, self.in_view.buf, self.in_view.C)
                self._detect(compact=False, fused_stem=True)
        elif kind == "emit_u8":  # inference_batches(): final arrays (classes ascending) inside the same graph
            def fn():
                if self.stem is None:
                    pack_u8(self.in_u8, self.in_view.buf, self.in_view.C)
                self._detect(compact=False, fused_stem=True, emit=True)
        elif kind == "gather_u8":  # emit_u8 + flat y3_cand records for the multi-GPU gather (one collective)
            def fn():
                if self.stem is None:
                    pack_u8(self.in_u8, self.in_view.buf, self.in_view.C)
                self._detect(compact=True, fused_stem=True, emit=True)
        elif kind == "det_f32":
            def fn():
                pack_f32(self.in_f32, self.in_view.buf, self.in_view.C)
                self._detect()
        else:
            raise KeyError(kind)
        return fn

    def launch(self, key):
        """Run program ``key`` on the current stream (CUDA-graph replay after the first call).
        ``key`` = ``(kind,)`` or ``(kind, prob_thresh, iou_thresh)``; thresholds are not part of the
        graph (they live in the device record), so every setting replays the same graph."""
        if self.dry:
            raise RuntimeError("this plan was built on the meta device and cannot run")
        if len(key) == 3:
            with torch.cuda.device(self.device):
                self.set_thresholds(key[1], key[2])
        key = key[:1]
        ent = self._graphs.get(key)
        if ent is None:
            fn = self._program(key)
            with torch.cuda.device(self.device):
                _lib.reset_launch_count()
                pdl_before = _lib.set_pdl(self.pdl)
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    fn()  # warm-up: sets kernel attributes, resolves driver entry points
                torch.cuda.current_stream(self.device).wait_stream(side)
                launches = _lib.launch_count() + (0 if key[0].startswith("dense") else 1)  # + counts.zero_()
                graph = None
                if self.use_graphs:
                    torch.cuda.synchronize(self.device)
                    graph = torch.cuda.CUDAGraph()
                    # thread_local: a CUDA call made by ANOTHER thread during the capture (ProcessGroupNCCL's
                    # watchdog polling events, a data-loader thread pinning memory) must not invalidate it
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        fn()
                _lib.set_pdl(pdl_before)
            ent = (fn, graph, launches)
            self._graphs[key] = ent
        fn, graph, _ = ent
        if graph is not None:
            graph.replay()
        else:
            fn()

    def launches(self, key):
        """Kernels launched by one run of program ``key`` (bench.py's gpu_launches)."""
        return self._graphs[key[:1]][2]

    # -- Darknet.forward ---------------------------------------------------------------------
    def forward_dense(self, x):
        with torch.cuda.device(self.device):
            self.in_f32.copy_(x.to(torch.float32), non_blocking=True)
            self.launch(("dense_f32",))
            return {"bbox_xywh": self.bbox.clone(), "class_prob": self.prob.clone(), "class_idx": self.cidx.clone()}

    # -- inference ------------------------------------------------------------------------------
    def detect(self, prob_thresh, iou_thresh, kind="det_u8"):
        """Run the fused detection program on whatever the static input buffer holds; results
        stay on the device: ``dets`` (flat y3_cand records of all images, image after image),
        ``det_counts`` [B], ``first_box`` [B, classes]."""
        with torch.cuda.device(self.device):
            self.launch((kind, float(prob_thresh), float(iou_thresh)))
        return self.dets, self.det_counts, self.first_box

    def memory_bound_ops(self):
        """(name, launch closure, algorithmic HBM bytes) of every memory-bound kernel of this plan: the
        pooling / elementwise block kernels that were not fused away, the input packing kernels and the
        dense YOLO decode of ``Darknet.forward`` (bench.py's ``hbm_roofline``)."""
        ops = list(self.mem_ops)
        px = self.B * self.H * self.W
        cin = self.net.net_info["channels"]
        pack_f32 = _lib.im2col3x3_nchw_f32 if self.first_im2col else _lib.pack_nchw_f32
        ops.append((pack_f32.__name__, lambda: pack_f32(self.in_f32, self.in_view.buf, self.in_view.C),
                    px * (cin * 4 + self.in_view.C * 2)))
        if self.in_u8 is not None and self.stem is None:
            pack_u8 = _lib.im2col3x3_bgr_u8 if self.first_im2col else _lib.pack_bgr_u8
            ops.append((pack_u8.__name__, lambda: pack_u8(self.in_u8, self.in_view.buf, self.in_view.C),
                        px * (3 + self.in_view.C * 2)))
        for hk, (d, logits) in enumerate(self.head_descs):
            boxes = d.n * d.num_anchors * d.g_h * d.g_w
            fields = d.num_anchors * (5 + d.num_classes)
            ops.append((f"yolo_decode_dense[{d.g_h}x{d.g_w}]",
                        lambda d=d, logits=logits: _lib.yolo_decode_dense(d, logits, self.bbox, self.prob, self.cidx),
                        d.n * d.g_h * d.g_w * fields * 4 + boxes * (16 + 4 + 8)))
        return ops

    def time_convs_in_sequence(self, passes=5):
        """Device time of every convolution launch measured INSIDE one in-order pass over the network:
        CUDA events are recorded between consecutive launches, so every kernel finds the L2 in the state
        its real predecessor left it in (``time_convs`` replays one launch back to back, which keeps small
        layers' operands L2-warm and flatters them).  A spin kernel in front gives the host a head start so
        the device never waits for a launch.  Returns (seconds per forward summed over convs — median pass,
        [(block, seconds, flops)])."""
        with torch.cuda.device(self.device):
            stream = torch.cuda.Stream(device=self.device)
            n = len(self.conv_ops)
            rows = []
            with torch.cuda.stream(stream):
                for _ in range(passes + 1):  # first pass = warm-up
                    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
                    torch.cuda._sleep(20_000_000)  # ~10 ms
                    evs[0].record(stream)
                    for j, (_, fn, _) in enumerate(self.conv_ops):
                        fn()
                        evs[j + 1].record(stream)
                    stream.synchronize()
                    rows.append([evs[j].elapsed_time(evs[j + 1]) * 1e-3 for j in range(n)])
            rows = rows[1:]
            order = sorted(range(len(rows)), key=lambda r: sum(rows[r]))
            med = rows[order[len(order) // 2]]
            per = [(blk, med[j], flops) for j, (blk, _, flops) in enumerate(self.conv_ops)]
            return sum(med), per

    def time_convs(self, iters=10):
        """Device time of every convolution launch, each timed ALONE: the launch is captured
        `iters` times into a small CUDA graph (so host launch latency does not pace the
        measurement) and the replay is bracketed by CUDA events on the same stream.
        Returns (seconds per forward summed over convs, [(block, seconds, flops)])."""
        with torch.cuda.device(self.device):
            per = []
            stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(stream):
                for blk, fn, flops in self.conv_ops:
                    fn()
                    stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                        for _ in range(iters):
                            fn()
                    g.replay()
                    stream.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    g.replay()
                    e1.record(stream)
                    stream.synchronize()
                    per.append((blk, e0.elapsed_time(e1) * 1e-3 / iters, flops))
                    del g
            return sum(p[1] for p in per), per


def records_to_numpy(rec):
    """int32 [K,8] y3_cand records -> (tlbr int64 [K,4], prob float32 [K], cls int64 [K], box int64 [K])."""
    rec = np.ascontiguousarray(rec)
    tlbr = rec[:, 0:4].astype(np.int64)
    prob = rec[:, 4].copy().view(np.float32)
    return tlbr, prob, rec[:, 5].astype(np.int64), rec[:, 6].astype(np.int64)
