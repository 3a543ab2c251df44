"""ctypes wrapper of oracle/nms_oracle.c (TEST INFRASTRUCTURE — not product code)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnms_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.nms_oracle.restype = ctypes.c_int64
        _lib.nms_oracle.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p]
    return _lib


def nms(tlbr, prob, class_idx=None, iou_thresh=0.3):
    """Same contract as postprocess_oracle.nms (kept indices, reference visiting order)."""
    tlbr = np.ascontiguousarray(tlbr, dtype=np.int64)
    prob = np.ascontiguousarray(prob, dtype=np.float32)
    n = tlbr.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int32)
    if class_idx is None:
        k = _load().nms_oracle(tlbr.ctypes.data, prob.ctypes.data, None, n, None, 0, float(iou_thresh),
                               keep.ctypes.data)
    else:
        class_idx = np.ascontiguousarray(class_idx, dtype=np.int64)
        order = np.asarray([int(c) for c in set(class_idx)], dtype=np.int64)  # Python set order (F6)
        k = _load().nms_oracle(tlbr.ctypes.data, prob.ctypes.data, class_idx.ctypes.data, n,
                               order.ctypes.data, order.size, float(iou_thresh), keep.ctypes.data)
    if k < 0:
        raise MemoryError("nms_oracle")
    return keep[:k].tolist()
