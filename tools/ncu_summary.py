"""Turns the ncu artefacts a gpurun call brought back into the text summaries kept under profiles/.
    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.md
    python tools/ncu_summary.py full gpurun_out/prof_fprop_r01.ncu-rep > profiles/r01_fprop_full.md"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn])
        v = float(r[mv].replace(",", ""))
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        if r[mn] == "gpu__time_duration.sum":
            a["n"] += 1
            a["us"] += v / 1e3 if r[mu] in ("ns", "nsecond") else (v * 1e3 if r[mu] in ("ms", "msecond") else v)
        elif r[mn] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[mu], 1.0)
            a["rd" if "read" in r[mn] else "wr"] += v * scale
    tot = sum(v["us"] for v in agg.values())
    n = sum(v["n"] for v in agg.values())
    print(f"# ncu launch list: {n} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"| `{k[:100]}` | {v['n']} | {v['us']:.1f} | {100 * v['us'] / tot:.1f}% | {v['rd'] / 1e6:.1f} | {v['wr'] / 1e6:.1f} |")
    return agg


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    print(f"# ncu --set full: {path}\n")
    for r in rows[2:]:
        print(f"## {r[h.index('Kernel Name')][:120]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in h:
                print(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
