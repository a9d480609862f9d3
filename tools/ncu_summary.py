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


FAMILIES = {"conv3x3_fprop": ("conv_fprop_kernel", "conv_fprop_halo_kernel", "conv_fprop_halo2_kernel", "conv_fprop_tr64_kernel",
                              "conv_fprop_tr128_kernel"),
            "conv3x3_wgrad": ("conv_wgrad_kernel", "conv_wgrad_rs64_kernel")}


def launches(path, traffic_out=None):
    """Summarises ONE training step of the launch list: the launches between the last two im2col3x3 kernels (one per
    forward pass), so warm-up steps and the end-to-end loop of bench.py do not dilute the shares. With a second path:
    also writes the per-launch DRAM traffic of the conv kernel families (bench.py's roofline.traffic)."""
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    idc, kn, mn, mv, mu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    per = collections.OrderedDict()  # launch id -> dict
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        d = per.setdefault(r[idc], {"name": re.sub(r"\(.*", "", r[kn]), "us": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r[mv].replace(",", ""))
        if r[mn] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if r[mu] in ("ns", "nsecond") else (v * 1e3 if r[mu] in ("ms", "msecond") else v)
        elif r[mn] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[mu], 1.0)
            d["rd" if "read" in r[mn] else "wr"] = v * scale
    seq = list(per.values())
    marks = [i for i, d in enumerate(seq) if "im2col3x3" in d["name"]]
    if len(marks) >= 2:
        seq = seq[marks[-2]:marks[-1]]
    agg = collections.OrderedDict()
    for d in seq:
        a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        for k in ("us", "rd", "wr"):
            a[k] += d[k]
    tot = sum(v["us"] for v in agg.values())
    n = sum(v["n"] for v in agg.values())
    print(f"# ncu launch list of one training step: {n} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, "
          f"serialised: compare shares)\n")
    print("| kernel | launches | total us | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"| `{k[:100]}` | {v['n']} | {v['us']:.1f} | {100 * v['us'] / tot:.1f}% | {v['rd'] / 1e6:.1f} | {v['wr'] / 1e6:.1f} |")
    if traffic_out:
        import json
        out = {}
        for fam, names in FAMILIES.items():
            sel = [v for k, v in agg.items() if any(re.search(r"\b" + nm + r"\b", k) for nm in names)]
            cnt = sum(v["n"] for v in sel)
            if cnt:
                out[fam] = {"launches": cnt, "dram_bytes_per_launch": sum(v["rd"] + v["wr"] for v in sel) / cnt,
                            "share_of_step_ncu": round(sum(v["us"] for v in sel) / tot, 4),
                            "source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over "
                                      f"the launches of one UNet 16x3x360x480 step ({path})"}
        json.dump(out, open(traffic_out, "w"), indent=1)
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
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
