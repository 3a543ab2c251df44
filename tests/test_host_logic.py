"""CPU: host-side logic of the package (parser, module tree, weight loader, execution plan,
result ordering), the C-ABI library's exports, and the no-fallback guarantees."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import yolov3_b200
from yolov3_b200 import _lib
from yolov3_b200.engine import Engine
from yolov3_b200.inference import _destinations, _order_like_reference, _set_order
from conftest import GOLDEN, MODELS, PKG, ROOT


def test_parse_config_and_cache_set_match_reference():
    g = json.load(open(os.path.join(GOLDEN, "parse_config.json")))
    for name, path in (("yolov3", f"{MODELS}/yolov3.cfg"), ("yolov3-tiny", f"{MODELS}/yolov3-tiny.cfg"),
                       ("yolov3-spp", f"{MODELS}/yolov3-spp.cfg"), ("micro", f"{GOLDEN}/micro.cfg")):
        net = yolov3_b200.Darknet(path, device="cuda")
        assert net.blocks == g[name]["blocks"]
        assert net.net_info == g[name]["net_info"]
        assert sorted(net.blocks_to_cache) == g[name]["blocks_to_cache"]


def test_parse_config_quirks(tmp_path):
    cfg = tmp_path / "q.cfg"
    cfg.write_text("[net]\nwidth=32\n# comment\n  \nheight=32\nchannels=3\nsteps=1,2\n\n[convolutional]\nfilters=16\n"
                   "size=3\nstride=1\npad=1\nactivation=linear\n[route]\nlayers=-1\n[yolo]\nmask=0\n"
                   "anchors=1,2, 3,4\nclasses=1\nscale=.5\n")
    blocks, net_info = yolov3_b200.parse_config(str(cfg))
    assert net_info == {"type": "net", "width": 32, "height": 32, "channels": 3, "steps": [1, 2]}
    assert blocks[1]["layers"] == [-1]            # scalar route -> list
    assert blocks[2]["mask"] == 0                 # single value stays scalar
    assert blocks[2]["anchors"] == [[1, 2], [3, 4]]
    assert blocks[2]["scale"] == 0.5 and blocks[0]["activation"] == "linear"


def test_module_tree_names_and_state_dict():
    net = yolov3_b200.Darknet(f"{MODELS}/yolov3-tiny.cfg", device="cuda")
    assert isinstance(net, torch.nn.Module) and len(net.modules_) == 24
    names = [n for n, _ in net.modules_[0].named_children()]
    assert names == ["conv_0", "batch_norm_0", "leaky_0"]
    assert [n for n, _ in net.modules_[15].named_children()] == ["conv_15"]  # linear head: no activation module
    assert net.modules_[15][0].bias is not None and net.modules_[0][0].bias is None
    assert [n for n, _ in net.modules_[11].named_children()] == ["maxpool_11"]
    assert "modules_.0.conv_0.weight" in net.state_dict()
    n_float = sum(p.numel() for p in net.parameters()) + sum(
        b.numel() for n, b in net.named_buffers() if "running" in n)
    assert n_float == 8858734  # yolov3-tiny.weights payload


def test_load_weights_fills_parameters_and_validates_length(tmp_path):
    net = yolov3_b200.Darknet(f"{GOLDEN}/micro.cfg", device="cuda")
    assert net.load_weights(f"{GOLDEN}/micro.weights") is net
    assert net.header.tolist() == [0, 2, 0, 0, 0]
    raw = np.fromfile(f"{GOLDEN}/micro.weights", dtype=np.float32)[5:]
    bn = net.modules_[0][1]
    assert np.array_equal(bn.bias.detach().numpy(), raw[:16])
    assert np.array_equal(bn.weight.detach().numpy(), raw[16:32])
    assert np.array_equal(bn.running_mean.numpy(), raw[32:48])
    assert np.array_equal(bn.running_var.numpy(), raw[48:64])
    assert np.array_equal(net.modules_[0][0].weight.detach().numpy().ravel(), raw[64:64 + 16 * 3 * 9])
    data = open(f"{GOLDEN}/micro.weights", "rb").read()
    short = tmp_path / "short.weights"
    short.write_bytes(data[: len(data) - 400])
    with pytest.raises(RuntimeError):
        yolov3_b200.Darknet(f"{GOLDEN}/micro.cfg", device="cuda").load_weights(str(short))
    longer = tmp_path / "long.weights"
    longer.write_bytes(data + b"\0" * 64)  # trailing values are ignored, as in the reference
    yolov3_b200.Darknet(f"{GOLDEN}/micro.cfg", device="cuda").load_weights(str(longer))


@pytest.mark.parametrize("name,size,convs,shortcuts,upsamples,M,gflop", [
    ("yolov3", 416, 75, 23, 2, 10647, 65.864), ("yolov3-tiny", 416, 13, 0, 1, 2535, 5.565),
    ("yolov3-spp", 608, 76, 23, 2, 22743, 141.449)])
def test_execution_plan_fuses_everything(name, size, convs, shortcuts, upsamples, M, gflop):
    net = yolov3_b200.Darknet(f"{MODELS}/{name}.cfg", device="cuda")
    eng = Engine(net, 2, size, size, torch.device("meta"))
    kinds = [re.sub(r"[\d_]+$", "", n) for n in eng.op_names]
    # every conv block is covered exactly once per program: a residual chain stands for two convs,
    # the uint8 stem is the alternative form of blocks 0-1 (which stay as convs for float32 input)
    assert kinds.count("conv") + 2 * kinds.count("chain") == convs
    if name != "yolov3-tiny":
        assert kinds.count("chain") == 1 and kinds.count("stem") == 1 and eng.stem is not None
        assert eng.op_groups.count("stem_unfused") == 2 and eng.op_groups.count("stem_fused") == 1
        assert len(eng.conv_ops) == convs - 2  # stem + chain each time two blocks as one launch
    else:
        assert eng.stem is None and "chain" not in kinds
    assert "add" not in kinds and "upsample" not in kinds and "copy" not in kinds  # all fused / zero-copy
    assert eng.num_fused["shortcut"] == shortcuts and eng.num_fused["upsample"] == upsamples
    assert eng.M == M and eng.num_classes == 80
    assert abs(eng.conv_flops / 2 / 1e9 - gflop) < 1e-3  # algorithmic FLOPs of SURVEY.md §8d
    if name == "yolov3-spp":
        assert kinds.count("spp") == 1 and "maxpool" not in kinds


def test_plan_falls_back_to_standalone_kernels_when_intermediate_is_shared(tmp_path):
    # the conv feeding the shortcut is ALSO routed: the add cannot be fused into its epilogue
    cfg = tmp_path / "shared.cfg"
    cfg.write_text(
        "[net]\nwidth=32\nheight=32\nchannels=3\n"
        "[convolutional]\nbatch_normalize=1\nfilters=16\nsize=3\nstride=1\npad=1\nactivation=leaky\n"
        "[convolutional]\nbatch_normalize=1\nfilters=16\nsize=3\nstride=1\npad=1\nactivation=leaky\n"
        "[shortcut]\nfrom=-2\nactivation=linear\n"
        "[route]\nlayers=-2,-1\n"
        "[upsample]\nstride=2\n"
        "[convolutional]\nfilters=18\nsize=1\nstride=1\npad=1\nactivation=linear\n"
        "[yolo]\nmask=0,1,2\nanchors=1,2, 3,4, 5,6\nclasses=1\n")
    net = yolov3_b200.Darknet(str(cfg), device="cuda")
    eng = Engine(net, 1, 32, 32, torch.device("meta"))
    kinds = [re.sub(r"[\d_]+$", "", n) for n in eng.op_names]
    assert "add" in kinds and "upsample" in kinds
    assert eng.num_fused["shortcut"] == 0


def test_unsupported_cfg_raises_instead_of_falling_back(tmp_path):
    cfg = tmp_path / "bad.cfg"
    cfg.write_text("[net]\nwidth=32\nheight=32\nchannels=3\n[convolutional]\nfilters=16\nsize=5\nstride=1\npad=1\n"
                   "activation=leaky\n[yolo]\nmask=0,1\nanchors=1,2, 3,4\nclasses=3\n")
    net = yolov3_b200.Darknet(str(cfg), device="cuda")
    with pytest.raises(NotImplementedError):
        Engine(net, 1, 32, 32, torch.device("meta"))


def test_cxywh_to_tlbr_known_answer():
    # tests/test_inference.py:12-23 of the reference
    z = np.load(os.path.join(GOLDEN, "nms.npz"))
    assert np.array_equal(yolov3_b200.cxywh_to_tlbr(z["kat_in"]), z["kat_out"])


def test_reference_class_visiting_order_is_reproduced():
    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(1, 400))
        classes = int(rng.choice([1, 2, 5, 9, 33, 80, 257, 1000]))
        cls = rng.integers(0, classes, n).astype(np.int64)
        first = np.full(max(classes, 1), np.iinfo(np.int32).max, dtype=np.int32)
        for pos, c in enumerate(cls):
            first[c] = min(first[c], pos)
        assert list(_set_order(first)) == list(set(cls))  # the reference's loop order (inference.py:247-250)
        # full permutation: records sorted (class asc, prob desc) -> reference order
        prob = rng.permutation(n).astype(np.float32)
        order = np.lexsort((-prob, cls))
        perm = _order_like_reference(cls[order], first)
        want = []
        for c in set(cls):
            members = np.where(cls == c)[0]
            want.extend(members[np.argsort(prob[members])[::-1]].tolist())
        assert order[perm].tolist() == want


def test_destinations_follow_the_set_order_of_every_image():
    """`inference` tells the device where each (image, class) group of kept detections goes; the
    layout must be image after image, class groups in the reference's set() order (ascending fast
    path for >= 19 present classes, real set emulation below that)."""
    rng = np.random.default_rng(1)
    for trial in range(60):
        B, C = int(rng.integers(1, 9)), int(rng.choice([3, 20, 80, 200]))
        kept = np.zeros((B, C), np.int32)
        first = np.full((B, C), np.iinfo(np.int32).max, np.int32)
        want = np.full((B, C), -1, np.int64)
        pos = 0
        for i in range(B):
            n = int(rng.integers(0, 300))
            cls = rng.integers(0, int(rng.integers(1, C + 1)), n).astype(np.int64)
            for j, c in enumerate(cls):
                first[i, c] = min(first[i, c], j)
            for c in set(cls):  # the reference's loop (inference.py:247-250)
                kept[i, c] = int(rng.integers(1, 1 + (cls == c).sum()))
                want[i, c] = pos
                pos += kept[i, c]
        dst, per_image = _destinations(kept, first)
        assert per_image.tolist() == kept.sum(1).tolist()
        present = kept > 0
        assert np.array_equal(dst[present], want[present])


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "yolov3_b200.h")).read()
    declared = set(re.findall(r"\b(y3_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.y3_abi_version() == _lib.ABI_VERSION == 6
    lib.y3_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.y3_last_error(), bytes)


def test_no_cpu_fallback():
    net = yolov3_b200.Darknet(f"{GOLDEN}/micro.cfg", device="cpu").load_weights(f"{GOLDEN}/micro.weights").eval()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU"):
            net.forward(torch.rand(1, 3, 64, 64))
        with pytest.raises(RuntimeError):
            yolov3_b200.inference(net, np.zeros((64, 64, 3), np.uint8), device="cuda")
        with pytest.raises(RuntimeError):
            yolov3_b200.non_max_suppression(np.zeros((2, 4), np.int64), np.ones(2, np.float32))
    with pytest.raises(RuntimeError):
        _lib.maxpool(0, 0, 1, 1, 1, 8, 8, 8, 2, 2) if torch.cuda.is_available() else _lib.require_device("cpu")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package or the C sources may touch it."""
    bad = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|nms_oracle", text, re.M):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
    # bench.py / __graft_entry__.py may use it only in the sanctioned places
    bench = open(os.path.join(ROOT, "bench.py")).read()
    for m in re.finditer(r"^.*\boracle\b.*$", bench, re.M):
        line = m.group(0)
        assert "baseline" in line or "reference" in line or line.lstrip().startswith(("#", '"', "from oracle", "import oracle")), line


def test_u8_scale_by_reciprocal_is_exact_after_bf16_rounding():
    """The im2col packing kernel multiplies by fl(1/255) instead of dividing by 255
    (yolov3/inference.py:333): identical bf16 result for every possible byte."""
    v = torch.arange(256, dtype=torch.float32)
    assert torch.equal((v / 255.0).bfloat16(), (v * np.float32(0.003921568859368563)).bfloat16())
    assert np.float32(0.003921568859368563) == np.float32(1.0) / np.float32(255.0)


def test_sub_batch_spans_cover_the_batch_in_order(monkeypatch):
    """inference() pipelines contiguous spans through the GPU: whatever the layout (default growing
    spans, forced equal spans, explicit split), the spans tile [0, batch) without gaps or empties."""
    from yolov3_b200.inference import _sub_batches
    for env in ({}, {"Y3_SUB_BATCHES": "4"}, {"Y3_SUB_BATCHES": "1"}, {"Y3_SUB_SPLIT": "1,3,4"}, {"Y3_SUB_SPLIT": "5,1,1,9"}):
        monkeypatch.delenv("Y3_SUB_BATCHES", raising=False)
        monkeypatch.delenv("Y3_SUB_SPLIT", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for batch in list(range(1, 70)) + [100, 127, 128, 256]:
            spans = _sub_batches(batch)
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(hi > lo for lo, hi in spans)
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    monkeypatch.delenv("Y3_SUB_SPLIT", raising=False)
    assert _sub_batches(64) == [(0, 8), (8, 32), (32, 64)]  # small first upload, the rest in two large plans
    assert _sub_batches(8) == [(0, 8)]


def test_stage_images_equals_np_stack_for_any_thread_count():
    """y3_stage_images (host threads of the library copying 256 KB chunks that may straddle images)
    is np.stack, for odd image sizes, more threads than chunks and repeated use of the parked pool."""
    rng = np.random.default_rng(5)
    for shape, n in (((37, 53, 3), 5), ((416, 416, 3), 16), ((600, 701, 3), 3), ((1, 1, 3), 2)):
        imgs = [rng.integers(0, 256, shape, dtype=np.uint8) for _ in range(n)]
        want = np.stack(imgs)
        for threads in (1, 3, 16, 64):
            dst = np.zeros_like(want)
            _lib.stage_images(dst, imgs, threads)
            assert np.array_equal(dst, want)


def test_reciprocal_division_is_exact_over_the_kernels_ranges():
    """The convolution kernels replace x / d by umulhi(x << 16, ceil(2^48 / d)) (conv_umma.cu fast_div,
    conv_patch.cu pt_div).  Exact whenever x * d < 2^48: checked here on the divisors the plans use
    (feature-map sizes, tiles per image, n tiles) at the largest x they see, plus random pairs."""
    def fast_div(x, d):
        m = ((1 << 48) + d - 1) // d
        return ((x << 16) * m) >> 64
    rng = np.random.default_rng(11)
    divisors = [1, 2, 3, 4, 8, 11, 13, 15, 19, 26, 28, 38, 40, 44, 52, 54, 76, 104, 106, 152, 169, 208, 304, 361, 416,
                608, 676, 1444, 2704, 5776, 10816, 23104, 43264, 92416, 173056, 369664]
    for d in divisors:
        xs = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 31) - 1 if d < (1 << 17) else (1 << 28)]
        xs += [int(v) for v in rng.integers(0, min(1 << 31, (1 << 48) // d), 2000)]
        for x in xs:
            if x * d < (1 << 48):
                assert fast_div(x, d) == x // d, (x, d)


def test_set_pdl_returns_the_previous_setting():
    """y3_set_pdl is host state only (no CUDA call): usable on a CPU box, returns what was set before."""
    first = _lib.set_pdl(False)
    try:
        assert _lib.set_pdl(True) is False
        assert _lib.set_pdl(True) is True
    finally:
        _lib.set_pdl(first)


def test_set_order_reordering_matches_python_sets():
    """_reorder_to_set_order on synthetic groups: any class subset, against a real Python set."""
    from yolov3_b200.inference import _reorder_to_set_order
    rng = np.random.default_rng(3)
    for _ in range(200):
        C = 80
        n_cls = int(rng.integers(1, 25))
        classes = rng.choice(C, n_cls, replace=False)
        cand_cls = rng.permutation(np.repeat(classes, rng.integers(1, 5, n_cls)))  # candidate order
        first = np.full(C, np.iinfo(np.int32).max, np.int32)
        kept = np.zeros(C, np.int32)
        for j, c in enumerate(cand_cls):
            first[c] = min(first[c], j)
        for c in classes:
            kept[c] = rng.integers(1, 4)
        asc = np.concatenate([np.full(kept[c], c) for c in range(C)])
        res = [np.stack([asc] * 4, 1).astype(np.int64), asc.astype(np.float32), asc.astype(np.int64)]
        out = _reorder_to_set_order(res, kept, first)
        want = np.concatenate([np.full(kept[c], c) for c in set(np.int64(c) for c in cand_cls)])
        assert np.array_equal(out[2], want)
