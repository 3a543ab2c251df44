"""Hot spots of an `ncu --page source --csv` dump: per kernel, the SASS instructions holding the most
warp-stall samples with their two dominant stall reasons, plus the kernel-wide stall mix.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_hot.py src.csv [min_share]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for k in kernels:
        hdr, data = k["hdr"], k["data"]
        iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
        stalls = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[iS] or 0) for r in data)
        print(f"=== {k['name'][:90]}: {len(data)} instructions, {tot} samples, "
              f"{sum(int(r[iEx] or 0) for r in data)} warp instructions")
        agg = {}
        for idx, r in enumerate(data):
            s = int(r[iS] or 0)
            for h, j in stalls:
                agg[h] = agg.get(h, 0) + int(r[j] or 0)
            if s > tot * thresh:
                top = sorted(((int(r[j] or 0), h) for h, j in stalls), reverse=True)[:2]
                print(f"{idx:5d} {100 * s / tot:5.1f}% ex={r[iEx]:>9} {r[iSrc][:72]:72s} {top}")
        print("stall mix:", [(h, round(100 * v / max(tot, 1), 1)) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])


if __name__ == "__main__":
    main()
