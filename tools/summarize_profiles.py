"""Turn the raw ncu artefacts a gpurun call brings back into the small summaries kept under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches.csv profiles/r01_launches_step.csv
    python tools/summarize_profiles.py raw gpurun_out/prof_conv.ncu-rep profiles/r01_conv_ncu_full.csv

`launches`: the `--metrics gpu__time_duration.sum` list -> every launch of the LAST step (from its
pack/im2col kernel on) with grid/block/µs, followed by a per-kernel aggregate with the share of the
step.  `raw`: `ncu -i … --page raw --csv` reduced to the columns the roofline argument uses.
"""
import csv
import collections
import re
import subprocess
import sys

RAW_COLS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]


def short(name):
    name = name.replace("void ", "").replace("y3::", "")
    m = re.match(r"([\w:]+(<[^(]*>)?)", name)
    return (m.group(1) if m else name)[:100]


def read_csv_after_header(path):
    lines = open(path).read().splitlines()
    i0 = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    return list(csv.DictReader(lines[i0:]))


def launches(src, dst):
    """src: ncu --csv log with gpu__time_duration.sum (and optionally dram__bytes_read.sum /
    dram__bytes_write.sum) per launch.  Also writes <dst stem>_traffic.json = DRAM bytes of the
    convolution launches of the step (bench.py's roofline.traffic)."""
    rows = read_csv_after_header(src)
    by_id = collections.OrderedDict()
    for r in rows:
        e = by_id.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        elif r["Metric Name"].startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            e[r["Metric Name"].split(".")[0].replace("dram__bytes_", "")] = val * scale
    recs = list(by_id.values())
    starts = [i for i, r in enumerate(recs) if "im2col" in r["name"] or "pack" in r["name"]
              or r["name"].startswith("conv_chain_kernel<(int)0>") or r["name"].startswith("conv_chain_kernel<0>")]
    step = recs[starts[-1]:]
    total = sum(r["us"] for r in step)
    has_dram = all("read" in r and "write" in r for r in step)
    agg = collections.OrderedDict()
    for r in step:
        a = agg.setdefault(r["name"], [0.0, 0, 0.0])
        a[0] += r["us"]
        a[1] += 1
        a[2] += (r.get("read", 0) + r.get("write", 0)) / 1e6
    with open(dst, "w") as f:
        f.write(f"# last step of {src}: {len(step)} launches, {total:.1f} us serialised (ncu, cold cache)\n")
        f.write("# --- per-kernel aggregate: kernel,launches,us,share_of_step,dram_MB\n")
        for n, (us, c, mb) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{n},{c},{us:.1f},{us / total:.4f},{mb:.1f}\n")
        f.write("# --- every launch: idx,kernel,grid,block,us,dram_read_MB,dram_write_MB\n")
        for i, r in enumerate(step):
            f.write(f"{i},{r['name']},\"{r['grid']}\",\"{r['block']}\",{r['us']:.2f},"
                    f"{r.get('read', 0) / 1e6:.2f},{r.get('write', 0) / 1e6:.2f}\n")
    if has_dram:
        import json
        conv = [r for r in step if r["name"].startswith(("conv_umma_kernel", "conv_chain_kernel", "conv_patch_kernel"))]
        out = {"source": dst, "conv_launches": len(conv),
               "dram_bytes_per_step": sum(r["read"] + r["write"] for r in conv),
               "conv_us_serialised": sum(r["us"] for r in conv), "step_us_serialised": total}
        json.dump(out, open(dst.rsplit(".", 1)[0] + "_traffic.json", "w"), indent=1)
        print(out)
    print(open(dst).read().split("# --- every")[0])


def raw(src, dst):
    """src: an .ncu-rep, or the `ncu -i rep --page raw --csv` export of one (reports of a whole step
    exceed what gpurun brings back, so the GPU box exports the raw page itself)."""
    if src.endswith(".csv"):
        txt = open(src).read()
    else:
        txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    while rows and "Kernel Name" not in rows[0]:
        rows = rows[1:]
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [hdr.index(c) for c in RAW_COLS if c in hdr]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for d in data:
            w.writerow([short(d[i]) if hdr[i] == "Kernel Name" else d[i] for i in cols])
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
