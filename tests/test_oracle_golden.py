"""CPU: the oracle restatement (oracle/) against the golden vectors generated from the LIVE
reference by tests/golden/make_golden.py.  Everything here is bit-exact except float tensors
that went through multi-threaded oneDNN reductions (tolerance stated where used)."""
import json
import os

import numpy as np
import torch

from oracle import darknet_oracle as DO
from oracle import nms_c
from oracle import postprocess_oracle as PO
from conftest import GOLDEN, MODELS


def test_parse_config_matches_reference_dump():
    g = json.load(open(os.path.join(GOLDEN, "parse_config.json")))
    for name, path in (("yolov3", f"{MODELS}/yolov3.cfg"), ("yolov3-tiny", f"{MODELS}/yolov3-tiny.cfg"),
                       ("yolov3-spp", f"{MODELS}/yolov3-spp.cfg"), ("micro", f"{GOLDEN}/micro.cfg")):
        blocks, net_info = DO.parse_config(path)
        keep = DO.resolve_routes(blocks)
        assert blocks == g[name]["blocks"]
        assert net_info == g[name]["net_info"]
        assert sorted(keep) == g[name]["blocks_to_cache"]
    assert len(g["yolov3"]["blocks"]) == 107 and len(g["yolov3-tiny"]["blocks"]) == 24
    assert len(g["yolov3-spp"]["blocks"]) == 114


def test_conv_blocks():
    z = np.load(os.path.join(GOLDEN, "conv_blocks.npz"))
    n = 0
    while f"c{n}_meta" in z:
        cin, cout, k, s, bn, leaky, H = z[f"c{n}_meta"]
        block = {"type": "convolutional", "filters": int(cout), "size": int(k), "stride": int(s), "pad": 1,
                 "activation": "leaky" if leaky else "linear"}
        if bn:
            block["batch_normalize"] = 1
        prm = {key: torch.from_numpy(z[f"c{n}_{key}"]) for key in
               ("weight", "bias", "bn_weight", "bn_bias", "bn_mean", "bn_var") if f"c{n}_{key}" in z}
        y = DO.conv_block(torch.from_numpy(z[f"c{n}_x"]), block, prm)
        # oneDNN may pick a different blocking on another host: 1e-5 of the tensor's range
        ref = torch.from_numpy(z[f"c{n}_y"])
        assert (y - ref).abs().max() <= 1e-5 * ref.abs().max()
        n += 1
    assert n == 6


def test_maxpool_zero_right_bottom_padding():
    z = np.load(os.path.join(GOLDEN, "maxpool.npz"))
    n = 0
    while f"p{n}_meta" in z:
        k, s = z[f"p{n}_meta"]
        y = DO.maxpool_block(torch.from_numpy(z[f"p{n}_x"]), {"size": int(k), "stride": int(s)})
        assert torch.equal(y, torch.from_numpy(z[f"p{n}_y"]))
        n += 1
    assert n == 6


def test_yolo_decode():
    z = np.load(os.path.join(GOLDEN, "yolo_decode.npz"))
    anchors = z["anchors"].tolist()
    n = 0
    while f"y{n}_x" in z:
        b, p, i = DO.yolo_decode(torch.from_numpy(z[f"y{n}_x"]), [anchors[m] for m in z[f"y{n}_mask"]])
        assert np.allclose(b.numpy(), z[f"y{n}_bbox"], rtol=1e-6, atol=0)
        assert np.allclose(p.numpy(), z[f"y{n}_prob"], rtol=1e-6, atol=0)
        assert np.array_equal(i.numpy(), z[f"y{n}_idx"])
        n += 1
    assert n == 4


def test_micro_network_forward_and_weights_roundtrip(tmp_path):
    blocks, net_info = DO.load_model(os.path.join(GOLDEN, "micro.cfg"))
    header, params = DO.read_weights(os.path.join(GOLDEN, "micro.weights"), blocks, net_info)
    assert header.tolist() == [0, 2, 0, 0, 0]
    # writer/reader round trip is byte-identical
    out = tmp_path / "rt.weights"
    DO.write_weights(str(out), params, blocks, net_info)
    assert out.read_bytes() == open(os.path.join(GOLDEN, "micro.weights"), "rb").read()
    z = np.load(os.path.join(GOLDEN, "micro_forward.npz"))
    cap = {}
    with torch.no_grad():
        o = DO.forward(torch.from_numpy(z["x"]), blocks, net_info, params, capture=cap)
    for k in z.files:
        if k.startswith("block"):
            ref = z[k]
            got = cap[int(k[5:])].numpy()
            assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max(), k
    assert np.allclose(o["bbox_xywh"].numpy(), z["bbox_xywh"], rtol=1e-4, atol=1e-6)
    assert np.allclose(o["class_prob"].numpy(), z["class_prob"], rtol=1e-4, atol=1e-7)
    assert (o["class_idx"].numpy() != z["class_idx"]).mean() < 1e-3


def test_weights_short_file_raises(tmp_path):
    blocks, net_info = DO.load_model(os.path.join(GOLDEN, "micro.cfg"))
    data = open(os.path.join(GOLDEN, "micro.weights"), "rb").read()
    short = tmp_path / "short.weights"
    short.write_bytes(data[:len(data) // 2])
    try:
        DO.read_weights(str(short), blocks, net_info)
    except RuntimeError:
        return
    raise AssertionError("short weights file must raise")


def test_postprocess_matches_reference_inference():
    z = np.load(os.path.join(GOLDEN, "postprocess.npz"))
    n = 0
    while f"q{n}_meta" in z:
        B, M, classes, H, W = z[f"q{n}_meta"]
        pt, it = z[f"q{n}_thr"]
        res = PO.postprocess(z[f"q{n}_bbox_xywh"].copy(), z[f"q{n}_class_prob"], z[f"q{n}_class_idx"],
                             [(int(H), int(W), 3)] * int(B), pt, it)
        for i, r in enumerate(res):
            assert np.array_equal(r[0], z[f"q{n}_img{i}_tlbr"]) and r[0].dtype == np.int64
            assert np.array_equal(r[1], z[f"q{n}_img{i}_prob"]) and r[1].dtype == np.float32
            assert np.array_equal(r[2], z[f"q{n}_img{i}_cls"]) and r[2].dtype == np.int64
        n += 1
    assert n == 4


def test_nms_numpy_and_c_oracles_match_reference_keep_lists():
    z = np.load(os.path.join(GOLDEN, "nms.npz"))
    for n in range(int(z["num_cases"][0])):
        tlbr, prob, cls = z[f"n{n}_tlbr"], z[f"n{n}_prob"], z[f"n{n}_cls"]
        per_class = bool(z[f"n{n}_meta"][2])
        thr = float(z[f"n{n}_thr"][0])
        want = z[f"n{n}_keep"].tolist()
        if tlbr.shape[0] == 0:
            assert want == []
            continue
        assert PO.nms(tlbr, prob, cls if per_class else None, thr) == want
        assert nms_c.nms(tlbr, prob, cls if per_class else None, thr) == want
    # the reference's own known-answer test (tests/test_inference.py:12-23)
    assert np.array_equal(PO.cxywh_to_tlbr(z["kat_in"]), z["kat_out"])


def test_preprocess():
    z = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    assert np.array_equal(PO.preprocess(list(z["images"])), z["inp"])


def test_micro_inference_end_to_end():
    blocks, net_info = DO.load_model(os.path.join(GOLDEN, "micro.cfg"))
    _, params = DO.read_weights(os.path.join(GOLDEN, "micro.weights"), blocks, net_info)
    z = np.load(os.path.join(GOLDEN, "micro_inference.npz"))
    imgs = list(z["images"])
    with torch.no_grad():
        o = DO.forward(torch.from_numpy(PO.preprocess(imgs)), blocks, net_info, params)
    res = PO.postprocess(o["bbox_xywh"].numpy(), o["class_prob"].numpy(), o["class_idx"].numpy(),
                         [im.shape for im in imgs], 0.3, 0.3)
    for i, r in enumerate(res):
        # float forward may differ in the last bits across hosts: compare detections as sets
        want = {tuple(t) + (int(c),) for t, c in zip(z[f"img{i}_tlbr"].tolist(), z[f"img{i}_cls"])}
        got = {tuple(t) + (int(c),) for t, c in zip(r[0].tolist(), r[2])}
        assert len(want ^ got) <= 0.02 * len(want)


def test_bf16_matched_oracle_equals_reference_modules_with_rounding_hooks():
    """oracle/bf16_matched.py vs the golden built from the reference's OWN modules (deep copy, BN folded
    into conv.weight, bf16 rounding hooks; tests/golden/make_golden.py::golden_bf16_matched)."""
    from oracle import bf16_matched as BM
    z = np.load(os.path.join(GOLDEN, "micro_bf16_matched.npz"))
    blocks, net_info = DO.load_model(os.path.join(GOLDEN, "micro.cfg"))
    _, params = DO.read_weights(os.path.join(GOLDEN, "micro.weights"), blocks, net_info)
    torch.set_num_threads(1)
    with torch.no_grad():
        o = BM.forward(torch.from_numpy(z["x"]), blocks, net_info, params)
    assert np.allclose(o["bbox_xywh"].numpy(), z["bbox_xywh"], rtol=1e-5, atol=1e-6)
    assert np.allclose(o["class_prob"].numpy(), z["class_prob"], rtol=1e-5, atol=1e-7)
    assert (o["class_idx"].numpy() != z["class_idx"]).mean() < 1e-3
    # and it is NOT the fp32 forward: rounding moves the heads measurably
    with torch.no_grad():
        f = DO.forward(torch.from_numpy(z["x"]), blocks, net_info, params)
    assert float((f["class_prob"] - o["class_prob"]).abs().max()) > 1e-4
