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
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn])
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: {n} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:100]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")


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
