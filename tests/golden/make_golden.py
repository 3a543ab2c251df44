"""Generate tests/golden/*.npz|json from the LIVE reference and pin the oracle to it.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

For every fixture the script (1) runs nrsyed/pytorch-yolov3's own code (imported, unmodified,
from /root/reference; harness-side shim ``np.int = int`` because the reference uses the alias
NumPy removed — yolov3/inference.py:353), (2) runs the oracle restatement (oracle/) on the same
inputs and ASSERTS agreement (bit-exact for every CPU path: same torch / numpy primitives in
the same order), and (3) stores inputs + reference outputs so tests/test_oracle_golden.py can
re-check the oracle anywhere, and the GPU parity tests can use the same vectors.
"""
import json
import os
import sys

import numpy as np

np.int = int  # harness shim for the reference

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("Y3_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import torch  # noqa: E402
import yolov3 as ref  # noqa: E402  (the reference package)
from yolov3 import darknet as ref_darknet  # noqa: E402
from yolov3 import inference as ref_inference  # noqa: E402

from oracle import darknet_oracle as DO  # noqa: E402
from oracle import postprocess_oracle as PO  # noqa: E402
from oracle import nms_c  # noqa: E402
from oracle import bf16_matched as BM  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(1)  # fixed reduction order for the stored float vectors


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    return o


def golden_parse_config():
    out = {}
    for name in ("yolov3-tiny", "yolov3", "yolov3-spp"):
        path = os.path.join(REF, "models", name + ".cfg")
        rb, rn = ref_darknet.parse_config(path)
        ob, on = DO.parse_config(path)
        assert rb == ob and rn == on, name
        # after Darknet.__init__ (absolute route indices, blocks_to_cache)
        net = ref.Darknet(path, device="cpu")
        ob2, _ = DO.parse_config(path)
        keep = DO.resolve_routes(ob2)
        assert net.blocks == ob2 and net.blocks_to_cache == keep, name
        out[name] = {"blocks": jsonable(net.blocks), "net_info": jsonable(rn),
                     "blocks_to_cache": sorted(net.blocks_to_cache)}
    # micro.cfg (ours) through the reference parser
    path = os.path.join(HERE, "micro.cfg")
    net = ref.Darknet(path, device="cpu")
    ob2, on = DO.parse_config(path)
    keep = DO.resolve_routes(ob2)
    assert net.blocks == ob2 and net.net_info == on and net.blocks_to_cache == keep
    out["micro"] = {"blocks": jsonable(net.blocks), "net_info": jsonable(on),
                    "blocks_to_cache": sorted(net.blocks_to_cache)}
    with open(os.path.join(HERE, "parse_config.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("parse_config: 4 cfgs identical")


def golden_micro_network():
    """Whole-network golden on micro.cfg: weights file -> reference load_weights -> forward."""
    cfg = os.path.join(HERE, "micro.cfg")
    blocks, net_info = DO.load_model(cfg)
    params = DO.synth_params(blocks, net_info, 64, seed=1234)
    wpath = os.path.join(HERE, "micro.weights")
    DO.write_weights(wpath, params, blocks, net_info)

    net = ref.Darknet(cfg, device="cpu").load_weights(wpath).eval()
    # weights round trip: what the oracle reads back == what the reference loaded
    _, rparams = DO.read_weights(wpath, blocks, net_info)
    for i, prm in rparams.items():
        conv = net.modules_[i][0]
        assert torch.equal(conv.weight.data, prm["weight"])
        if "bn_bias" in prm:
            bn = net.modules_[i][1]
            assert torch.equal(bn.bias.data, prm["bn_bias"]) and torch.equal(bn.weight.data, prm["bn_weight"])
            assert torch.equal(bn.running_mean, prm["bn_mean"]) and torch.equal(bn.running_var, prm["bn_var"])
        else:
            assert torch.equal(conv.bias.data, prm["bias"])

    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, 64, 64, generator=g)

    # per-block outputs of the reference via forward hooks on its own modules
    ref_blocks = {}
    hooks = []
    for i, (b, m) in enumerate(zip(net.blocks, net.modules_)):
        if b["type"] in ("convolutional", "maxpool", "upsample"):
            hooks.append(m.register_forward_hook(lambda mod, inp, out, i=i: ref_blocks.__setitem__(i, out.detach().clone())))
        if b["type"] == "yolo":
            hooks.append(m.register_forward_hook(
                lambda mod, inp, out, i=i: ref_blocks.__setitem__(f"head{i}", inp[0].detach().clone())))
    with torch.no_grad():
        rout = net.forward(x.clone())
    for h in hooks:
        h.remove()

    cap = {}
    with torch.no_grad():
        oout = DO.forward(x.clone(), blocks, net_info, rparams, capture=cap)
    for k in ("bbox_xywh", "class_prob", "class_idx"):
        assert torch.equal(rout[k], oout[k]), k
    for k, v in ref_blocks.items():
        assert torch.equal(v, cap[k]), k
    np.savez_compressed(
        os.path.join(HERE, "micro_forward.npz"), x=x.numpy(),
        bbox_xywh=rout["bbox_xywh"].numpy(), class_prob=rout["class_prob"].numpy(),
        class_idx=rout["class_idx"].numpy(),
        **{f"block{k}" if isinstance(k, int) else k: v.numpy() for k, v in ref_blocks.items()})
    print("micro network: forward + %d per-block tensors bit-identical" % len(ref_blocks))

    # inference() end to end on the micro net (uint8 BGR images, non-square original size)
    rng = np.random.default_rng(1234)
    imgs = [rng.integers(0, 256, (64, 64, 3), dtype=np.uint8) for _ in range(2)]
    with torch.no_grad():
        rres = ref.inference(net, [im.copy() for im in imgs], device="cpu", prob_thresh=0.3,
                             nms_iou_thresh=0.3, resize=False)
        inp = torch.from_numpy(PO.preprocess(imgs))
        o = DO.forward(inp, blocks, net_info, rparams)
    ores = PO.postprocess(o["bbox_xywh"].numpy(), o["class_prob"].numpy(), o["class_idx"].numpy(),
                          [im.shape for im in imgs], 0.3, 0.3)
    for r, q in zip(rres, ores):
        for a, b in zip(r, q):
            assert a.dtype == b.dtype and np.array_equal(a, b)
    np.savez_compressed(os.path.join(HERE, "micro_inference.npz"), images=np.stack(imgs),
                        **{f"img{i}_{n}": r[j] for i, r in enumerate(rres)
                           for j, n in enumerate(("tlbr", "prob", "cls"))})
    print("micro inference(): %s detections, identical" % [len(r[1]) for r in rres])


def golden_conv_blocks():
    """Teacher-forced conv block goldens from the reference's own module builder."""
    cases = [  # (cin, cout, k, stride, bn, leaky, H)
        (16, 32, 3, 1, True, True, 10), (32, 64, 3, 2, True, True, 12), (64, 32, 1, 1, True, True, 7),
        (128, 256, 3, 1, True, True, 6), (64, 21, 1, 1, False, False, 5), (48, 32, 3, 1, True, True, 9),
    ]
    store = {}
    g = torch.Generator().manual_seed(11)
    for n, (cin, cout, k, s, bn, leaky, H) in enumerate(cases):
        block = {"type": "convolutional", "filters": cout, "size": k, "stride": s, "pad": 1,
                 "activation": "leaky" if leaky else "linear"}
        if bn:
            block = {"type": "convolutional", "batch_normalize": 1, **{k_: v for k_, v in block.items() if k_ != "type"}}
        mods = ref_darknet.blocks2modules([block], {"channels": cin})
        seq = mods[0].eval()
        prm = {}
        conv = seq[0]
        conv.weight.data = torch.randn(conv.weight.shape, generator=g) * (2.0 / (cin * k * k)) ** 0.5
        prm["weight"] = conv.weight.data.clone()
        if bn:
            bnm = seq[1]
            bnm.weight.data = torch.rand(cout, generator=g) * 0.4 + 0.8
            bnm.bias.data = torch.randn(cout, generator=g) * 0.1
            bnm.running_mean = torch.randn(cout, generator=g) * 0.2
            bnm.running_var = torch.rand(cout, generator=g) + 0.5
            prm.update(bn_weight=bnm.weight.data.clone(), bn_bias=bnm.bias.data.clone(),
                       bn_mean=bnm.running_mean.clone(), bn_var=bnm.running_var.clone())
        else:
            conv.bias.data = torch.randn(cout, generator=g) * 0.5
            prm["bias"] = conv.bias.data.clone()
        x = torch.randn(2, cin, H, H, generator=g)
        with torch.no_grad():
            y = seq(x.clone())
            yo = DO.conv_block(x.clone(), block, prm)
        assert torch.equal(y, yo), n
        store[f"c{n}_meta"] = np.array([cin, cout, k, s, int(bn), int(leaky), H])
        store[f"c{n}_x"] = x.numpy()
        store[f"c{n}_y"] = y.numpy()
        for kk, v in prm.items():
            store[f"c{n}_{kk}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "conv_blocks.npz"), **store)
    print("conv blocks: %d cases bit-identical" % len(cases))


def golden_maxpool():
    store = {}
    g = torch.Generator().manual_seed(5)
    for n, (k, s, H, C) in enumerate([(2, 2, 8, 16), (2, 1, 7, 16), (5, 1, 9, 8), (9, 1, 9, 8), (13, 1, 9, 8),
                                      (2, 2, 13, 8)]):
        x = torch.randn(2, C, H, H, generator=g) - 0.7  # mostly negative: zero padding must win (F3)
        m = ref_darknet.MaxPool2d(kernel_size=k, stride=s)
        y = m(x.clone())
        yo = DO.maxpool_block(x.clone(), {"size": k, "stride": s})
        assert torch.equal(y, yo), n
        store[f"p{n}_meta"] = np.array([k, s])
        store[f"p{n}_x"] = x.numpy()
        store[f"p{n}_y"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "maxpool.npz"), **store)
    print("maxpool: identical (incl. zero right/bottom padding)")


def golden_yolo_decode():
    store = {}
    g = torch.Generator().manual_seed(3)
    all_anchors = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
    for n, (mask, gh, classes) in enumerate([([6, 7, 8], 3, 80), ([3, 4, 5], 5, 80), ([0, 1, 2], 4, 80),
                                             ([0, 1, 2], 6, 2)]):
        x = torch.randn(2, 3 * (5 + classes), gh, gh, generator=g) * 2.0
        layer = ref_darknet.YOLOLayer(all_anchors, mask, device="cpu")
        rb, rp, ri = layer(x.clone())
        ob, op, oi = DO.yolo_decode(x.clone(), [all_anchors[m] for m in mask])
        assert torch.equal(rb, ob) and torch.equal(rp, op) and torch.equal(ri, oi), n
        store[f"y{n}_mask"] = np.array(mask)
        store[f"y{n}_x"] = x.numpy()
        store[f"y{n}_bbox"] = rb.numpy()
        store[f"y{n}_prob"] = rp.numpy()
        store[f"y{n}_idx"] = ri.numpy()
    store["anchors"] = np.array(all_anchors)
    np.savez_compressed(os.path.join(HERE, "yolo_decode.npz"), **store)
    print("yolo decode: identical")


class _StubNet:
    """Stands in for Darknet so the reference's inference() post-processing can be driven with
    chosen decoded tensors (it only touches net.net_info and net.forward)."""

    def __init__(self, out, size):
        self.out, self.net_info = out, {"height": size, "width": size}

    def forward(self, inp):
        return {k: torch.from_numpy(v.copy()) for k, v in self.out.items()}


def synth_decoded(rng, B, M, classes):
    xy = rng.random((B, M, 2), dtype=np.float32)
    wh = (0.02 + 0.4 * rng.random((B, M, 2), dtype=np.float32)).astype(np.float32)
    # distinct probabilities per image: a permutation of M distinct float32 values (tie-free)
    base = np.linspace(0.011, 0.989, M, dtype=np.float32)
    assert np.unique(base).size == M
    prob = np.stack([rng.permutation(base) for _ in range(B)])
    idx = rng.integers(0, classes, (B, M)).astype(np.int64)
    return {"bbox_xywh": np.concatenate([xy, wh], axis=2), "class_prob": prob, "class_idx": idx}


def golden_bf16_matched():
    """The bf16-matched oracle (oracle/bf16_matched.py) against the same thing built from the
    reference's OWN modules (SURVEY.md §8c last row): deep copy of the reference net, BatchNorm folded
    into conv.weight (rounded to bf16) and turned into a pure bias, forward hooks rounding block
    outputs to bf16 at the plan's rounding points."""
    import copy
    cfg = os.path.join(HERE, "micro.cfg")
    blocks, net_info = DO.load_model(cfg)
    wpath = os.path.join(HERE, "micro.weights")
    net = copy.deepcopy(ref.Darknet(cfg, device="cpu").load_weights(wpath).eval())
    _, params = DO.read_weights(wpath, blocks, net_info)
    fused = BM.single_consumer_shortcuts(blocks)
    hooks = []
    for i, (b, m) in enumerate(zip(net.blocks, net.modules_)):
        if b["type"] != "convolutional":
            continue
        conv = m[0]
        W, bias = BM.fold_bn(params[i])
        with torch.no_grad():
            conv.weight.copy_(BM.bf16(W))
            if len(m) > 1 and isinstance(m[1], torch.nn.BatchNorm2d):
                bn = m[1]
                # gamma = 1, mean = 0, var = 1 - eps: (x - 0) / sqrt((1 - eps) + eps) * 1 + b' 
                bn.weight.fill_(1.0), bn.running_mean.zero_(), bn.running_var.fill_(1.0 - bn.eps)
                assert float(torch.sqrt(bn.running_var[0] + bn.eps)) == 1.0
                bn.bias.copy_(bias)
            else:
                conv.bias.copy_(bias)
        head = i + 1 < len(blocks) and blocks[i + 1]["type"] == "yolo"
        hooks.append(m.register_forward_pre_hook(lambda mod, inp: (BM.bf16(inp[0]),)))
        if not (head or i in fused):
            hooks.append(m.register_forward_hook(lambda mod, inp, out: BM.bf16(out)))
    # the reference adds shortcuts inline (darknet.py:376-379): round their result through a wrapper
    # around torch.Tensor.__add__ is not possible from outside, so the reference forward is replayed
    # block by block with ITS modules and ITS cache rule, rounding the shortcut sums.
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, 64, 64, generator=g)
    cached, outs = {}, []
    xx = BM.bf16(x.clone())
    with torch.no_grad():
        for i, b in enumerate(net.blocks):
            t = b["type"]
            if t in ("convolutional", "maxpool", "upsample"):
                xx = net.modules_[i](xx)
            elif t == "route":
                xx = torch.cat(tuple(cached[j] for j in b["layers"]), dim=1)
            elif t == "shortcut":
                xx = BM.bf16(cached[i - 1] + cached[i + b["from"]])
            elif t == "yolo":
                outs.append(net.modules_[i][0](xx))
            if i in net.blocks_to_cache:
                cached[i] = xx
        bbox = torch.cat([o[0] for o in outs], dim=1)
        bbox[:, :, 2:4] /= torch.tensor([net.net_info["width"], net.net_info["height"]])
        rprob, ridx = torch.cat([o[1] for o in outs], dim=1), torch.cat([o[2] for o in outs], dim=1)
        o = BM.forward(x.clone(), blocks, net_info, params)
    for h in hooks:
        h.remove()
    # BN as a pure bias goes through batch_norm's (x - 0) / sqrt(1) * 1 + b': exact
    assert torch.equal(o["bbox_xywh"], bbox) and torch.equal(o["class_prob"], rprob) and torch.equal(o["class_idx"], ridx)
    np.savez_compressed(os.path.join(HERE, "micro_bf16_matched.npz"), x=x.numpy(), bbox_xywh=bbox.numpy(),
                        class_prob=rprob.numpy(), class_idx=ridx.numpy())
    print("bf16-matched oracle == reference modules with fold + rounding hooks (micro.cfg): identical")


def golden_postprocess():
    rng = np.random.default_rng(99)
    store = {}
    cases = [(2, 400, 80, (48, 64), 0.05, 0.3), (1, 300, 5, (37, 53), 0.5, 0.45), (2, 200, 3, (64, 64), 0.0, 0.0),
             (1, 50, 80, (20, 20), 0.999, 0.3)]
    for n, (B, M, classes, hw, pt, it) in enumerate(cases):
        out = synth_decoded(rng, B, M, classes)
        imgs = [np.zeros((hw[0], hw[1], 3), np.uint8) for _ in range(B)]
        rres = ref.inference(_StubNet(out, 64), imgs, device="cpu", prob_thresh=pt, nms_iou_thresh=it, resize=False)
        ores = PO.postprocess(out["bbox_xywh"].copy(), out["class_prob"], out["class_idx"],
                              [im.shape for im in imgs], pt, it)
        for r, q in zip(rres, ores):
            for a, b in zip(r, q):
                assert a.dtype == b.dtype and np.array_equal(a, b), n
        store[f"q{n}_meta"] = np.array([B, M, classes, hw[0], hw[1]])
        store[f"q{n}_thr"] = np.array([pt, it], dtype=np.float64)
        for k, v in out.items():
            store[f"q{n}_{k}"] = v
        for i, r in enumerate(rres):
            store[f"q{n}_img{i}_tlbr"], store[f"q{n}_img{i}_prob"], store[f"q{n}_img{i}_cls"] = r
    np.savez_compressed(os.path.join(HERE, "postprocess.npz"), **store)
    print("inference() post-processing: %d cases identical" % len(cases))


def golden_nms():
    rng = np.random.default_rng(2024)
    store = {}
    n_case = 0
    for (n, classes, size, thr) in [(300, 5, 100, 0.3), (500, 80, 416, 0.3), (64, 1, 30, 0.5), (257, 7, 60, 0.0),
                                    (1, 3, 10, 0.3), (0, 3, 10, 0.3), (33, 2, 12, 0.9)]:
        cx = rng.integers(0, size, n)
        cy = rng.integers(0, size, n)
        w = rng.integers(0, size // 2 + 1, n)
        h = rng.integers(0, size // 2 + 1, n)
        tlbr = ref.cxywh_to_tlbr(np.stack([cx, cy, w, h], axis=1).astype(np.int64)) if n else np.zeros((0, 4), np.int64)
        assert np.array_equal(tlbr, PO.cxywh_to_tlbr(np.stack([cx, cy, w, h], axis=1).astype(np.int64))) if n else True
        prob = rng.permutation(np.linspace(0.01, 0.99, max(n, 1), dtype=np.float32))[:n]
        cls = rng.integers(0, classes, n).astype(np.int64)
        for per_class in (True, False):
            k_ref = ref.non_max_suppression(tlbr, prob, cls if per_class else None, thr) if n else []
            k_orc = PO.nms(tlbr, prob, cls if per_class else None, thr) if n else []
            k_c = nms_c.nms(tlbr, prob, cls if per_class else None, thr) if n else []
            assert [int(v) for v in k_ref] == k_orc == k_c, (n, classes, per_class)
            store[f"n{n_case}_tlbr"], store[f"n{n_case}_prob"], store[f"n{n_case}_cls"] = tlbr, prob, cls
            store[f"n{n_case}_meta"] = np.array([n, classes, int(per_class)])
            store[f"n{n_case}_thr"] = np.array([thr])
            store[f"n{n_case}_keep"] = np.array([int(v) for v in k_ref], dtype=np.int64)
            n_case += 1
    # exact-equality edge: two boxes whose IoU is exactly the threshold are BOTH kept (iou > thr)
    tlbr = np.array([[0, 0, 9, 9], [0, 0, 9, 4]], dtype=np.int64)  # inter 50, union 100 -> 0.5
    prob = np.array([0.9, 0.8], dtype=np.float32)
    for thr, expect in ((0.5, [0, 1]), (0.49999, [0])):
        k_ref = ref.non_max_suppression(tlbr, prob, None, thr)
        assert [int(v) for v in k_ref] == expect == PO.nms(tlbr, prob, None, thr) == nms_c.nms(tlbr, prob, None, thr)
        store[f"n{n_case}_tlbr"], store[f"n{n_case}_prob"], store[f"n{n_case}_cls"] = tlbr, prob, np.zeros(2, np.int64)
        store[f"n{n_case}_meta"] = np.array([2, 1, 0])
        store[f"n{n_case}_thr"] = np.array([thr])
        store[f"n{n_case}_keep"] = np.array(expect, dtype=np.int64)
        n_case += 1
    # the reference's own known-answer vector (tests/test_inference.py:12-23)
    kat_in = np.array([[5, 8, 10, 13, 10000], [100, 200, 30, 17, 19000]], dtype=np.int64)
    kat_out = np.array([[0, 2, 10, 14, 10000], [85, 192, 115, 208, 19000]])
    assert (ref.cxywh_to_tlbr(kat_in) == kat_out).all() and (PO.cxywh_to_tlbr(kat_in) == kat_out).all()
    store["kat_in"], store["kat_out"] = kat_in, kat_out
    store["num_cases"] = np.array([n_case])
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **store)
    print("nms: %d cases identical (reference == numpy oracle == C oracle)" % n_case)


def golden_preprocess():
    rng = np.random.default_rng(8)
    imgs = [rng.integers(0, 256, (6, 9, 3), dtype=np.uint8) for _ in range(2)]
    captured = {}

    class Net(_StubNet):
        def forward(self, inp):
            captured["inp"] = inp.numpy().copy()
            return super().forward(inp)

    out = synth_decoded(rng, 2, 10, 3)
    ref.inference(Net(out, 64), imgs, device="cpu", resize=False)
    assert np.array_equal(captured["inp"], PO.preprocess(imgs))
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), images=np.stack(imgs), inp=captured["inp"])
    print("preprocess: identical")


if __name__ == "__main__":
    nms_c.build()
    golden_parse_config()
    golden_conv_blocks()
    golden_maxpool()
    golden_yolo_decode()
    golden_micro_network()
    golden_bf16_matched()
    golden_postprocess()
    golden_nms()
    golden_preprocess()
    print("all goldens written to", HERE)
