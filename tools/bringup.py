"""GPU bring-up harness: runs each kernel check in its own subprocess (a trap or hang in one
kernel must not poison the rest) and prints one line per case.  Usage on the GPU box:

    python tools/bringup.py [--only conv] [--timeout 120]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-yolov3_b200"))

CONV_CASES = [
    # name, n, h, w, cin, cout, k, s, leaky, res, up, f32, force_im2col
    ("1x1_tiled_k64_n32", 2, 13, 13, 64, 32, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_tiled_k1024_n256", 4, 13, 13, 1024, 256, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_im2col_k64_n32", 2, 13, 13, 64, 32, 1, 1, 1, 0, 0, 0, 1),
    ("1x1_im2col_k1024_n256", 4, 13, 13, 1024, 256, 1, 1, 1, 0, 0, 0, 1),
    ("3x3_s1_c64_n128", 2, 26, 26, 64, 128, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_s2_c32_n64_sw64", 2, 32, 32, 32, 64, 3, 2, 1, 0, 0, 0, 0),
    ("3x3_s1_c16_n32_sw32", 1, 64, 64, 16, 32, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_s1_c16_n16_sw32", 1, 32, 32, 16, 16, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_s1_c48_n32_sw32", 2, 16, 16, 48, 32, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_s1_res", 2, 26, 26, 128, 256, 3, 1, 1, 1, 0, 0, 0),
    ("1x1_upsample", 2, 13, 13, 512, 256, 1, 1, 1, 0, 1, 0, 0),
    ("1x1_head_f32_linear", 2, 13, 13, 1024, 256, 1, 1, 0, 0, 0, 1, 0),
    ("3x3_s2_c64_n128", 2, 52, 52, 64, 128, 3, 2, 1, 0, 0, 0, 0),
    ("3x3_persistent_676tiles", 8, 104, 104, 64, 128, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_c512_n1024_13", 8, 13, 13, 512, 1024, 3, 1, 1, 0, 0, 0, 0),
    ("1x1_n512", 2, 26, 26, 256, 512, 1, 1, 1, 0, 0, 0, 0),
    ("3x3_pair_res_52", 4, 52, 52, 128, 256, 3, 1, 1, 1, 0, 0, 0),
    ("3x3_pair_s2", 4, 52, 52, 256, 512, 3, 2, 1, 0, 0, 0, 0),
    ("1x1_pair_tail", 3, 13, 13, 512, 256, 1, 1, 1, 0, 0, 0, 0),
    ("3x3_pair_many_tiles", 16, 52, 52, 128, 256, 3, 1, 1, 1, 0, 0, 0),
]


def run_conv_case(case):
    import torch
    import torch.nn.functional as F
    from yolov3_b200 import _lib
    name, n, h, w, cin, cout, k, s, leaky, res, up, f32, force = case
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    pad = (k - 1) // 2
    x = torch.randn(n, cin, h, w, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(dev).to(torch.bfloat16)
    bias = torch.randn(cout, generator=g).to(dev)
    ho = (h + 2 * pad - k) // s + 1
    wo = (w + 2 * pad - k) // s + 1
    r = torch.randn(n, cout, ho, wo, generator=g).to(dev).to(torch.bfloat16) if res else None
    ref = F.conv2d(x.float(), wt.float(), bias, stride=s, padding=pad)
    if leaky:
        ref = F.leaky_relu(ref, 0.1)
    if res:
        ref = ref + r.float()
    if up:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_krsc = wt.permute(0, 2, 3, 1).contiguous()
    r_nhwc = r.permute(0, 2, 3, 1).contiguous() if res else None
    oh, ow = (2 * ho, 2 * wo) if up else (ho, wo)
    out = None
    for direct, one_cta in ((False, False), (False, True), (True, True)):
        y = torch.full((n, oh, ow, cout), float("nan"), device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        t0 = time.time()
        _lib.conv2d(x_nhwc.data_ptr(), w_krsc, bias, y.data_ptr(), n=n, h=h, w_in=w, cin=cin, cout=cout, ksize=k,
                    stride=s, pad=pad, ld_x=cin, ld_y=cout, leaky=leaky, res_ptr=r_nhwc.data_ptr() if res else None,
                    ld_res=cout, out_f32=bool(f32), upsample2x=bool(up), force_im2col=bool(force),
                    force_direct=direct, force_1cta=one_cta)
        torch.cuda.synchronize()
        dt = time.time() - t0
        got = y.float().permute(0, 3, 1, 2)
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        nan = int(torch.isnan(got).sum().item())
        ok = nan == 0 and err < 1e-2
        cur = {"case": name + ("/direct" if direct else "/staged") + ("/1cta" if one_cta else "/auto"), "ok": ok, "rel_err": err, "nan": nan,
               "ms_first_call": dt * 1e3}
        if out is None or not ok:
            out = cur
        if not ok:
            break
    if not ok:  # locate the damage
        d = (got - ref).abs()
        bad = (d > 1e-2 * ref.abs().max()) | torch.isnan(got)
        idx = bad.nonzero()
        out["bad_count"] = int(bad.sum().item())
        out["first_bad"] = idx[:5].tolist()
        out["bad_by_channel_block16"] = [int(bad[:, c:c + 16].sum().item()) for c in range(0, min(cout, 64), 16)]
        out["bad_rows"] = sorted(set(idx[:, 2].tolist()))[:16]
        out["bad_cols"] = sorted(set(idx[:, 3].tolist()))[:16]
    return out


def run_pointwise():
    import torch
    import torch.nn.functional as F
    from yolov3_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    res = []

    def nhwc(t):
        return t.permute(0, 2, 3, 1).contiguous()

    for (k, s, H, C) in [(2, 2, 26, 64), (2, 1, 13, 512), (2, 2, 13, 16)]:
        x = (torch.randn(2, C, H, H, generator=g) - 0.7).to(dev).to(torch.bfloat16)
        xp = F.pad(x.float(), (0, k - 1, 0, k - 1)) if (k > 1 and s == 1) else x.float()
        ref = F.max_pool2d(xp, k, s)
        y = torch.empty(2, ref.shape[2], ref.shape[3], C, device=dev, dtype=torch.bfloat16)
        xn = nhwc(x)
        _lib.maxpool(xn.data_ptr(), y.data_ptr(), 2, H, H, C, C, C, k, s)
        torch.cuda.synchronize()
        res.append({"case": f"maxpool_k{k}s{s}", "ok": bool(torch.equal(y.float().permute(0, 3, 1, 2), ref))})
    # spp into a 4*C concat buffer
    C, H = 64, 19
    x = (torch.randn(2, C, H, H, generator=g) - 0.5).to(dev).to(torch.bfloat16)
    buf = torch.zeros(2, H, H, 4 * C, device=dev, dtype=torch.bfloat16)
    buf[..., 3 * C:] = nhwc(x)
    base = buf.data_ptr()
    _lib.spp3(base + 3 * C * 2, base + 2 * C * 2, base + 1 * C * 2, base, 2, H, H, C, 4 * C, 4 * C)
    torch.cuda.synchronize()
    refs = [F.max_pool2d(F.pad(x.float(), (0, k - 1, 0, k - 1)), k, 1) for k in (13, 9, 5)] + [x.float()]
    ref = torch.cat(refs, dim=1)
    res.append({"case": "spp3_concat", "ok": bool(torch.equal(buf.float().permute(0, 3, 1, 2), ref))})
    # add / copy / upsample
    a = torch.randn(2, 32, 9, 9, generator=g).to(dev).to(torch.bfloat16)
    b = torch.randn(2, 32, 9, 9, generator=g).to(dev).to(torch.bfloat16)
    y = torch.empty(2, 9, 9, 32, device=dev, dtype=torch.bfloat16)
    an, bn = nhwc(a), nhwc(b)  # keep the temporaries alive while the kernel reads them
    _lib.add(an.data_ptr(), bn.data_ptr(), y.data_ptr(), 2 * 81, 32, 32, 32, 32)
    torch.cuda.synchronize()
    res.append({"case": "add", "ok": bool(torch.equal(y, nhwc((a.float() + b.float()).to(torch.bfloat16))))})
    y = torch.zeros(2, 9, 9, 64, device=dev, dtype=torch.bfloat16)
    _lib.copy_channels(an.data_ptr(), y.data_ptr() + 32 * 2, 2 * 81, 32, 32, 64)
    torch.cuda.synchronize()
    res.append({"case": "copy_channels", "ok": bool(torch.equal(y[..., 32:], nhwc(a)) and (y[..., :32] == 0).all())})
    y = torch.empty(2, 18, 18, 32, device=dev, dtype=torch.bfloat16)
    _lib.upsample2x(an.data_ptr(), y.data_ptr(), 2, 9, 9, 32, 32, 32)
    torch.cuda.synchronize()
    res.append({"case": "upsample2x", "ok": bool(torch.equal(y.permute(0, 3, 1, 2), F.interpolate(a, scale_factor=2)))})
    # packing
    xf = torch.rand(2, 3, 10, 12, generator=g).to(dev)
    y = torch.empty(2, 10, 12, 16, device=dev, dtype=torch.bfloat16)
    _lib.pack_nchw_f32(xf, y, 16)
    torch.cuda.synchronize()
    ok = torch.equal(y[..., :3], nhwc(xf).to(torch.bfloat16)) and bool((y[..., 3:] == 0).all())
    res.append({"case": "pack_nchw_f32", "ok": bool(ok)})
    xu = torch.randint(0, 256, (2, 10, 12, 3), generator=g, dtype=torch.uint8).to(dev)
    _lib.pack_bgr_u8(xu, y, 16)
    torch.cuda.synchronize()
    ref = (xu.flip(3).float() / 255.0).to(torch.bfloat16)
    ok = torch.equal(y[..., :3], ref) and bool((y[..., 3:] == 0).all())
    res.append({"case": "pack_bgr_u8", "ok": bool(ok)})
    return res


def run_decode_nms():
    import numpy as np
    import torch
    from yolov3_b200 import _lib
    from oracle import darknet_oracle as DO
    from oracle import postprocess_oracle as PO
    from oracle import nms_c
    dev = torch.device("cuda:0")
    res = []
    g = torch.Generator().manual_seed(3)
    all_anchors = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
    B, classes = 3, 80
    grids = [(13, [6, 7, 8]), (26, [3, 4, 5]), (52, [0, 1, 2])]
    M = sum(3 * gg * gg for gg, _ in grids)
    bbox = torch.empty(B, M, 4, device=dev)
    prob = torch.empty(B, M, device=dev)
    idx = torch.empty(B, M, dtype=torch.int64, device=dev)
    cap = M
    cands = torch.zeros(B, cap, 8, dtype=torch.int32, device=dev)
    counts = torch.zeros(B, dtype=torch.int32, device=dev)
    orig_hw = torch.tensor([[375, 500], [416, 416], [480, 640]], dtype=torch.int32, device=dev)
    off = 0
    ob, op, oi = [], [], []
    thr = 0.05
    for gg, mask in grids:
        x = torch.randn(B, 3 * (5 + classes), gg, gg, generator=g) * 2.0
        anchors = [all_anchors[m] for m in mask]
        b_, p_, i_ = DO.yolo_decode(x.clone(), anchors)
        ob.append(b_), op.append(p_), oi.append(i_)
        logits = torch.zeros(B, gg, gg, 256)
        logits[..., :255] = x.permute(0, 2, 3, 1)
        logits = logits.to(dev).contiguous()
        d = _lib.make_head_desc(B, gg, gg, anchors, classes, 256, off, M, 608, 608)
        _lib.yolo_decode_dense(d, logits, bbox, prob, idx)
        _lib.yolo_decode_cands(d, logits, thr, orig_hw, cands, counts, cap)
        off += 3 * gg * gg
    torch.cuda.synchronize()
    ob = torch.cat(ob, 1)
    ob[:, :, 2:4] = ob[:, :, 2:4] / torch.tensor([608, 608])
    op, oi = torch.cat(op, 1), torch.cat(oi, 1)
    eb = ((bbox.cpu() - ob).abs() / ob.abs().clamp_min(1e-6)).max().item()
    ep = ((prob.cpu() - op).abs() / op.abs().clamp_min(1e-12)).max().item()
    ei = int((idx.cpu() != oi).sum().item())
    res.append({"case": "decode_dense", "ok": eb < 1e-5 and ep < 1e-5 and ei == 0, "bbox_rel": eb, "prob_rel": ep,
                "idx_mismatch": ei})
    # candidates: compare as sets keyed by box index against the oracle post-processing up to tlbr
    cnt = counts.cpu().numpy()
    cd = cands.cpu().numpy()
    mism, borderline = 0, 0
    for i in range(B):
        mask = op[i].numpy() >= thr
        box = ob[i].numpy()[mask].copy()
        hw = orig_hw[i].cpu().numpy()
        box[:, [0, 2]] *= hw[1]
        box[:, [1, 3]] *= hw[0]
        tlbr = PO.cxywh_to_tlbr(box.astype(np.int64))
        want = {int(m): (tuple(t), int(c)) for m, t, c in zip(np.nonzero(mask)[0], tlbr, oi[i].numpy()[mask])}
        got = {int(r[6]): ((int(r[0]), int(r[1]), int(r[2]), int(r[3])), int(r[5])) for r in cd[i, :cnt[i]]}
        for m in set(want) | set(got):
            if want.get(m) != got.get(m):
                if abs(float(op[i, m]) - thr) < 1e-6 or (m in want and m in got and
                                                         max(abs(a - b) for a, b in zip(want[m][0], got[m][0])) <= 1):
                    borderline += 1
                else:
                    mism += 1
    res.append({"case": "decode_cands", "ok": mism == 0, "mismatch": mism, "borderline_1px_or_thresh": borderline,
                "counts": cnt.tolist()})
    # NMS on the device candidates vs the C oracle on the same candidates
    ws = torch.empty(_lib.nms_workspace_bytes(B, cap, classes), dtype=torch.uint8, device=dev)
    for per_class in (1, 0):
        sorted_ = torch.zeros_like(cands)
        keep = torch.zeros(B, cap, dtype=torch.uint8, device=dev)
        first = torch.zeros(B, classes, dtype=torch.int32, device=dev)
        _lib.nms(cands, counts, B, cap, classes, 0.3, per_class, sorted_, keep, first, ws)
        dets = torch.zeros_like(cands)
        dcount = torch.zeros(B, dtype=torch.int32, device=dev)
        _lib.compact_kept(sorted_, keep, counts, B, cap, dets.view(-1, 8), dcount, 1)
        torch.cuda.synchronize()
        sd, kp, dc = sorted_.cpu().numpy(), keep.cpu().numpy(), dcount.cpu().numpy()
        bad = 0
        flat = dets.view(-1, 8).cpu().numpy()
        pos = 0
        for i in range(B):
            c = cd[i, :cnt[i]]
            tlbr = c[:, :4].astype(np.int64)
            pr = c[:, 4].copy().view(np.float32)
            cl = c[:, 5].astype(np.int64)
            kept = nms_c.nms(tlbr, pr, cl if per_class else None, 0.3)
            want = set(int(c[j, 6]) for j in kept)
            got = set(int(b) for b, k in zip(sd[i, :cnt[i], 6], kp[i, :cnt[i]]) if k)
            bad += len(want ^ got)
            got_flat = set(int(b) for b in flat[pos:pos + dc[i], 6])
            bad += len(want ^ got_flat)
            pos += dc[i]
        res.append({"case": f"nms_per_class{per_class}", "ok": bad == 0, "set_diff": bad, "kept": dc.tolist()})
    return res


def child(kind, idx):
    if kind == "conv":
        print("RESULT " + json.dumps(run_conv_case(CONV_CASES[idx])))
    elif kind == "pointwise":
        for r in run_pointwise():
            print("RESULT " + json.dumps(r))
    elif kind == "decode_nms":
        for r in run_decode_nms():
            print("RESULT " + json.dumps(r))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", nargs=2)
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=150)
    a = ap.parse_args()
    if a.child:
        child(a.child[0], int(a.child[1]))
        return
    jobs = [("conv", i) for i in range(len(CONV_CASES))] + [("pointwise", 0), ("decode_nms", 0)]
    if a.only:
        jobs = [j for j in jobs if j[0] == a.only]
    n_ok = n_bad = 0
    for kind, i in jobs:
        label = CONV_CASES[i][0] if kind == "conv" else kind
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", kind, str(i)],
                               capture_output=True, text=True, timeout=a.timeout)
            lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
            if not lines:
                print(f"FAIL {label}: rc={p.returncode} no result\n   stderr tail: {p.stderr[-600:]}")
                n_bad += 1
            for ln in lines:
                r = json.loads(ln[7:])
                print(("ok   " if r.get("ok") else "FAIL ") + json.dumps(r))
                n_ok += bool(r.get("ok"))
                n_bad += not r.get("ok")
        except subprocess.TimeoutExpired:
            print(f"FAIL {label}: TIMEOUT after {a.timeout}s")
            n_bad += 1
        sys.stdout.flush()
    print(f"bringup: {n_ok} ok, {n_bad} failed")


if __name__ == "__main__":
    main()
