"""ctypes binding of ``libyolov3_b200.so`` (C ABI declared in ``include/yolov3_b200.h``).

PyTorch is used for device memory and streams only: every wrapper takes torch
tensors, passes ``data_ptr()`` / shapes / the current CUDA stream through the
C ABI, and raises ``RuntimeError`` with ``y3_last_error()`` on a non-zero status.
There is no fallback: if the shared library is missing this module raises at
import of the first symbol (``lib()``), and every kernel wrapper refuses CPU
tensors.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_longlong, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# Y3_LIB: another build of the same ABI (A/B measurements of kernel changes on one box)
LIB_PATH = os.environ.get("Y3_LIB") or os.path.join(_HERE, "libyolov3_b200.so")

# Symbols include/yolov3_b200.h declares (tests check the .so exports exactly these).
EXPORTS = (
    "y3_abi_version", "y3_last_error", "y3_check_device", "y3_launch_count", "y3_reset_launch_count", "y3_set_pdl",
    "y3_stage_images", "y3_conv2d", "y3_conv2d_yolo_head", "y3_conv_chain_stem_u8", "y3_conv_chain_res64", "y3_maxpool", "y3_spp3", "y3_add", "y3_copy_channels", "y3_upsample2x",
    "y3_pack_nchw_f32", "y3_pack_bgr_u8", "y3_im2col3x3_nchw_f32", "y3_im2col3x3_bgr_u8", "y3_yolo_decode_dense", "y3_yolo_decode_cands",
    "y3_nms_workspace_bytes", "y3_nms", "y3_plan_destinations", "y3_compact_kept", "y3_emit_detections",
    "y3_debug_conv_trace", "y3_debug_set_trap_record",
)

ABI_VERSION = 6


class ConvDesc(ctypes.Structure):
    """``y3_conv_desc``."""
    _fields_ = [(n, c_int32) for n in (
        "n", "h", "w", "cin", "cout", "ksize", "stride", "pad", "ld_x", "ld_y", "ld_res",
        "leaky", "out_f32", "upsample2x", "flags")]


class ChainDesc(ctypes.Structure):
    """``y3_chain_desc``."""
    _fields_ = [(n, c_int32) for n in ("n", "h", "w", "ld_x", "ld_y", "leaky1", "leaky2")]


class HeadDesc(ctypes.Structure):
    """``y3_head_desc``."""
    _fields_ = [(n, c_int32) for n in (
        "n", "g_h", "g_w", "num_anchors", "num_classes", "ld", "box_offset", "boxes_per_image")] + [
        ("anchor_w", c_float * 8), ("anchor_h", c_float * 8), ("train_w", c_float), ("train_h", c_float)]


class Thresholds(ctypes.Structure):
    """``y3_thresholds`` (16 bytes; the device-resident copy is what the kernels read)."""
    _fields_ = [("prob_thresh", c_float), ("reserved_", c_float), ("iou_thresh", c_double)]


# numpy / torch view of ``y3_cand`` (32 bytes)
CAND_WORDS = 8  # int32 words per record: x1 y1 x2 y2 prob(f32 bits) cls box pad

_lib = None


def lib():
    """Load the shared library once; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C pytorch-yolov3_b200/csrc`). yolov3_b200 has no CPU / eager fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.y3_abi_version.restype = ctypes.c_int
    L.y3_last_error.restype = c_char_p
    L.y3_check_device.argtypes = [ctypes.c_int]
    L.y3_launch_count.restype = c_longlong
    L.y3_reset_launch_count.restype = None
    L.y3_stage_images.argtypes = [c_void_p, POINTER(c_void_p), c_int32, c_int64, c_int32]
    L.y3_conv2d.argtypes = [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.y3_conv2d_yolo_head.argtypes = [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, POINTER(HeadDesc), c_float,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    L.y3_conv_chain_stem_u8.argtypes = [POINTER(ChainDesc)] + [c_void_p] * 7
    L.y3_conv_chain_res64.argtypes = [POINTER(ChainDesc)] + [c_void_p] * 7
    L.y3_maxpool.argtypes = [c_void_p, c_void_p] + [c_int32] * 8 + [c_void_p]
    L.y3_spp3.argtypes = [c_void_p] * 4 + [c_int32] * 6 + [c_void_p]
    L.y3_add.argtypes = [c_void_p] * 3 + [c_int64] + [c_int32] * 4 + [c_void_p]
    L.y3_copy_channels.argtypes = [c_void_p] * 2 + [c_int64] + [c_int32] * 3 + [c_void_p]
    L.y3_upsample2x.argtypes = [c_void_p] * 2 + [c_int32] * 6 + [c_void_p]
    L.y3_pack_nchw_f32.argtypes = [c_void_p] * 2 + [c_int32] * 5 + [c_void_p]
    L.y3_pack_bgr_u8.argtypes = [c_void_p] * 2 + [c_int32] * 4 + [c_void_p]
    L.y3_im2col3x3_nchw_f32.argtypes = [c_void_p] * 2 + [c_int32] * 5 + [c_void_p]
    L.y3_im2col3x3_bgr_u8.argtypes = [c_void_p] * 2 + [c_int32] * 4 + [c_void_p]
    L.y3_yolo_decode_dense.argtypes = [POINTER(HeadDesc)] + [c_void_p] * 5
    L.y3_yolo_decode_cands.argtypes = [POINTER(HeadDesc), c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int32, c_void_p]
    L.y3_nms_workspace_bytes.argtypes = [c_int32] * 3
    L.y3_nms_workspace_bytes.restype = c_size_t
    L.y3_nms.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_double, c_void_p, c_int32, c_void_p, c_void_p,
                         c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.y3_plan_destinations.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    L.y3_debug_conv_trace.argtypes = [c_void_p, ctypes.c_int]
    L.y3_debug_set_trap_record.argtypes = [c_void_p]
    L.y3_emit_detections.argtypes = [c_void_p] * 4 + [c_int32] * 3 + [c_void_p] * 4
    L.y3_compact_kept.argtypes = [c_void_p] * 3 + [c_int32] * 2 + [c_void_p] * 2 + [c_int32, c_void_p]
    if L.y3_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libyolov3_b200.so ABI {L.y3_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"libyolov3_b200: {lib().y3_last_error().decode()} (status {rc})")


_checked_devices = set()


def require_device(device):
    """Raise unless ``device`` is a CUDA sm_100 device usable by the library."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("yolov3_b200 runs on CUDA sm_100a devices only (no CPU fallback); got device "
                           f"'{device}'")
    if not torch.cuda.is_available():
        raise RuntimeError("yolov3_b200 needs a CUDA device (no CPU fallback) and none is available")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        _check(lib().y3_check_device(idx))
        _checked_devices.add(idx)
    return torch.device("cuda", idx)


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("yolov3_b200 kernels take CUDA tensors only")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().y3_launch_count())


def reset_launch_count():
    lib().y3_reset_launch_count()


def set_pdl(on):
    """Programmatic dependent launch for subsequent launches / graph captures; returns the previous setting."""
    return bool(lib().y3_set_pdl(1 if on else 0))


_trap_records = {}


def enable_trap_record(device):
    """Register a pinned host buffer as the watchdog record of ``device`` (see y3_debug_set_trap_record);
    returns it (int64 tensor of 8 words; word 0 != 0 after a watchdog trap)."""
    device = torch.device(device)
    rec = _trap_records.get(device.index)
    if rec is None:
        rec = torch.zeros(12, dtype=torch.int64).pin_memory()
        with torch.cuda.device(device):
            _check(lib().y3_debug_set_trap_record(rec.data_ptr()))
        _trap_records[device.index] = rec
    return rec


def describe_trap_record(rec):
    """Human-readable form of a watchdog record, or None when no trap was recorded."""
    w = [int(v) & 0xFFFFFFFFFFFFFFFF for v in rec.tolist()]
    if w[0] == 0:
        return None
    files = {1: "conv_umma.cu", 2: "conv_patch.cu", 3: "conv_chain.cu"}

    def one(v):
        return (f"{files.get(v[0] >> 32, '?')}:{v[0] & 0xFFFFFFFF} block {v[1] & 0xFFFFFFFF} of {v[3] & 0xFFFFFFFF} "
                f"thread {v[1] >> 32} of {v[3] >> 32} barrier 0x{v[2] & 0xFFFFFFFF:x} parity {v[2] >> 32}")
    return f"mbarrier watchdog: {w[5]} threads timed out; first: {one(w[8:12])}; last: {one(w[1:5])}"


def stage_images(dst, images, threads):
    """``np.stack(images)`` into ``dst`` (host array, e.g. a pinned staging buffer) by the library's
    host threads; ``images`` are equally shaped C-contiguous uint8 arrays."""
    n = len(images)
    ptrs = (c_void_p * n)(*[im.ctypes.data for im in images])
    _check(lib().y3_stage_images(dst.ctypes.data, ptrs, n, images[0].nbytes, threads))


# ---- kernels --------------------------------------------------------------------------------

def conv2d(x_ptr, w, bias, y_ptr, *, n, h, w_in, cin, cout, ksize, stride, pad, ld_x, ld_y, leaky,
           res_ptr=None, ld_res=0, out_f32=False, upsample2x=False, force_im2col=False, force_direct=False, force_1cta=False,
           force_stream_weights=False):
    """Raw-pointer form used by the engine plan (views into concat buffers are plain pointers)."""
    d = ConvDesc(n, h, w_in, cin, cout, ksize, stride, pad, ld_x, ld_y, ld_res, int(leaky), int(out_f32),
                 int(upsample2x), (1 if force_im2col else 0) | (2 if force_direct else 0) | (4 if force_1cta else 0)
                 | (8 if force_stream_weights else 0))
    _check(lib().y3_conv2d(ctypes.byref(d), x_ptr, _ptr(w), _ptr(bias), res_ptr, y_ptr, _stream()))


def conv2d_yolo_head(x_ptr, w, bias, head, prob_thresh, orig_hw, cands, counts, cap, *, n, h, w_in, cin, ld_x,
                     dev_thresholds=None):
    """1x1 YOLO head convolution with decode + threshold + candidate append fused into its epilogue.
    ``dev_thresholds``: device tensor holding a ``y3_thresholds`` record read at run time (overrides
    ``prob_thresh``)."""
    d = ConvDesc(n, h, w_in, cin, 256, 1, 1, 0, ld_x, 256, 0, 0, 0, 0, 0)
    _check(lib().y3_conv2d_yolo_head(ctypes.byref(d), x_ptr, _ptr(w), _ptr(bias), ctypes.byref(head),
                                     float(prob_thresh), _ptr(dev_thresholds), _ptr(orig_hw), _ptr(cands),
                                     _ptr(counts), cap, _stream()))


def conv_chain_stem_u8(img, w1, b1, w2, b2, y_ptr, *, ld_y, leaky1=True, leaky2=True):
    """uint8 BGR images [N,H,W,3] -> conv3x3(3->32) -> conv3x3/2(32->64), one kernel (conv_chain.cu)."""
    n, h, w, c = img.shape
    assert c == 3 and img.dtype == torch.uint8
    d = ChainDesc(n, h, w, 0, ld_y, int(leaky1), int(leaky2))
    _check(lib().y3_conv_chain_stem_u8(ctypes.byref(d), _ptr(img), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), y_ptr,
                                       _stream()))


def conv_chain_res64(x_ptr, w1, b1, w2, b2, y_ptr, *, n, h, w, ld_x, ld_y, leaky1=True, leaky2=True):
    """Residual unit x -> conv1x1(64->32) -> conv3x3(32->64) + x, one kernel (conv_chain.cu)."""
    d = ChainDesc(n, h, w, ld_x, ld_y, int(leaky1), int(leaky2))
    _check(lib().y3_conv_chain_res64(ctypes.byref(d), x_ptr, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), y_ptr,
                                     _stream()))


def maxpool(x_ptr, y_ptr, n, h, w, c, ld_x, ld_y, ksize, stride):
    _check(lib().y3_maxpool(x_ptr, y_ptr, n, h, w, c, ld_x, ld_y, ksize, stride, _stream()))


def spp3(x_ptr, y5_ptr, y9_ptr, y13_ptr, n, h, w, c, ld_x, ld_y):
    _check(lib().y3_spp3(x_ptr, y5_ptr, y9_ptr, y13_ptr, n, h, w, c, ld_x, ld_y, _stream()))


def add(a_ptr, b_ptr, y_ptr, pixels, c, ld_a, ld_b, ld_y):
    _check(lib().y3_add(a_ptr, b_ptr, y_ptr, pixels, c, ld_a, ld_b, ld_y, _stream()))


def copy_channels(x_ptr, y_ptr, pixels, c, ld_x, ld_y):
    _check(lib().y3_copy_channels(x_ptr, y_ptr, pixels, c, ld_x, ld_y, _stream()))


def upsample2x(x_ptr, y_ptr, n, h, w, c, ld_x, ld_y):
    _check(lib().y3_upsample2x(x_ptr, y_ptr, n, h, w, c, ld_x, ld_y, _stream()))


def pack_nchw_f32(x, y, c_pad):
    n, c, h, w = x.shape
    _check(lib().y3_pack_nchw_f32(_ptr(x), _ptr(y), n, c, h, w, c_pad, _stream()))


def pack_bgr_u8(x, y, c_pad):
    n, h, w, c = x.shape
    assert c == 3
    _check(lib().y3_pack_bgr_u8(_ptr(x), _ptr(y), n, h, w, c_pad, _stream()))


def im2col3x3_nchw_f32(x, y, k_pad):
    n, c, h, w = x.shape
    _check(lib().y3_im2col3x3_nchw_f32(_ptr(x), _ptr(y), n, c, h, w, k_pad, _stream()))


def im2col3x3_bgr_u8(x, y, k_pad):
    n, h, w, c = x.shape
    assert c == 3
    _check(lib().y3_im2col3x3_bgr_u8(_ptr(x), _ptr(y), n, h, w, k_pad, _stream()))


def make_head_desc(n, g_h, g_w, anchors, num_classes, ld, box_offset, boxes_per_image, train_w, train_h):
    d = HeadDesc()
    d.n, d.g_h, d.g_w = n, g_h, g_w
    d.num_anchors, d.num_classes, d.ld = len(anchors), num_classes, ld
    d.box_offset, d.boxes_per_image = box_offset, boxes_per_image
    for i, (aw, ah) in enumerate(anchors):
        d.anchor_w[i] = float(aw)
        d.anchor_h[i] = float(ah)
    d.train_w, d.train_h = float(train_w), float(train_h)
    return d


def yolo_decode_dense(desc, logits, bbox_xywh, class_prob, class_idx):
    _check(lib().y3_yolo_decode_dense(ctypes.byref(desc), _ptr(logits), _ptr(bbox_xywh), _ptr(class_prob),
                                      _ptr(class_idx), _stream()))


def yolo_decode_cands(desc, logits, prob_thresh, orig_hw, cands, counts, cap, dev_thresholds=None):
    _check(lib().y3_yolo_decode_cands(ctypes.byref(desc), _ptr(logits), float(prob_thresh), _ptr(dev_thresholds),
                                      _ptr(orig_hw), _ptr(cands), _ptr(counts), cap, _stream()))


def nms_workspace_bytes(n, cap, num_classes):
    return int(lib().y3_nms_workspace_bytes(n, cap, num_classes))


def nms(cands, counts, n, cap, num_classes, iou_thresh, per_class, sorted_out, keep, class_first_box, workspace,
        class_start=None, class_kept=None, dev_thresholds=None):
    _check(lib().y3_nms(_ptr(cands), _ptr(counts), n, cap, num_classes, float(iou_thresh), _ptr(dev_thresholds),
                        int(per_class),
                        _ptr(sorted_out), _ptr(keep), _ptr(class_first_box), _ptr(class_start), _ptr(class_kept),
                        _ptr(workspace), workspace.numel() * workspace.element_size(), _stream()))


def plan_destinations(class_kept, n, num_segments, dst_off, det_counts):
    """Ascending-class destinations + per-image totals (``det_counts`` has n + 1 entries: last = total)."""
    _check(lib().y3_plan_destinations(_ptr(class_kept), n, num_segments, _ptr(dst_off), _ptr(det_counts), _stream()))


def emit_detections(sorted_in, keep, class_start, dst_off, n, cap, num_segments, tlbr, prob, cls):
    _check(lib().y3_emit_detections(_ptr(sorted_in), _ptr(keep), _ptr(class_start), _ptr(dst_off), n, cap,
                                    num_segments, _ptr(tlbr), _ptr(prob), _ptr(cls), _stream()))


def compact_kept(sorted_in, keep, counts, n, cap, dets, det_counts, flat):
    _check(lib().y3_compact_kept(_ptr(sorted_in), _ptr(keep), _ptr(counts), n, cap, _ptr(dets), _ptr(det_counts),
                                 int(flat), _stream()))
