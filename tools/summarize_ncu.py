"""Summarise ncu outputs into small text files under profiles/ (run in the build container, no GPU needed).

  python tools/summarize_ncu.py launches gpurun_out/r1a_launches.csv profiles/r1_launches_summary.md
  python tools/summarize_ncu.py full gpurun_out/r1a_prof_gemm.ncu-rep profiles/r1_gemm_full.md
  python tools/summarize_ncu.py gemm_traffic gpurun_out/r2e_gemm_dram.csv gpurun_out/r2e_gemm_shapes.json profiles/gemm_traffic.json
  python tools/summarize_ncu.py hbm gpurun_out/r2e_prof_misc.ncu-rep profiles/r2e_hbm_kernels.md   (GB/s from dram bytes / duration)
"""
import json
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def launches(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<.*", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        agg[name][0] += 1
        agg[name][1] += v * scale
        total += v * scale
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| {name} | {n} | {ms:.3f} | {100 * ms / total:.1f}% | {1000 * ms / n:.1f} |\n")
        f.write(f"\ntotal {total:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")
    print(open(out).read()[:3000])


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput",
        "sm__pipe_tensor", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor",
        "l1tex__data_bank_conflicts", "lts__t_sector_hit_rate", "sm__cycles_elapsed.avg.per_second",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active", "dram__cycles_active"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr = rd[0]
    units = rd[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({rep})\n\n")
        for row in rd[2:]:
            d = dict(zip(hdr, row))
            f.write(f"## {d.get('Kernel Name', '?')[:120]}  grid={d.get('Grid Size')} block={d.get('Block Size')}\n\n")
            for h, u, v in zip(hdr, units, row):
                if any(k in h for k in KEYS):
                    f.write(f"- {h} [{u}] = {v}\n")
            f.write("\n")
    print(open(out).read()[:4000])


def _metric_rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(lines))))


def _to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def _to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)


def gemm_traffic(csv_path, shapes_path, out):
    """Zip the per-launch ncu rows of the GEMM kernels (launch order) with tools/gemm_shape_trace.py's shape list."""
    per_launch = defaultdict(dict)
    for r in _metric_rows(csv_path):
        i = int(r["ID"])
        per_launch[i]["name"] = r["Kernel Name"]
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            per_launch[i][m] = _to_bytes(r["Metric Value"], r["Metric Unit"])
        elif m == "gpu__time_duration.sum":
            per_launch[i]["us"] = _to_us(r["Metric Value"], r["Metric Unit"])
    rows = [per_launch[i] for i in sorted(per_launch)]
    shapes = json.load(open(shapes_path))
    if len(rows) != len(shapes):
        raise SystemExit(f"{len(rows)} GEMM launches under ncu vs {len(shapes)} traced: orders cannot be matched")
    agg = defaultdict(lambda: {"launches": 0, "dram_bytes": 0.0, "algorithmic_bytes": 0.0, "us": 0.0, "flops": 0.0})
    for r, sh in zip(rows, shapes):
        a = agg[sh["key"]]
        a["launches"] += 1
        a["dram_bytes"] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
        a["algorithmic_bytes"] += sh["algorithmic_bytes"]
        a["us"] += r.get("us", 0.0)
        a["flops"] += sh["flops"]
    by_shape = {}
    for k, a in agg.items():
        n = a["launches"]
        by_shape[k] = {"launches": n, "dram_bytes_per_launch": a["dram_bytes"] / n,
                       "algorithmic_bytes_per_launch": a["algorithmic_bytes"] / n,
                       "traffic_over_algorithmic": a["dram_bytes"] / a["algorithmic_bytes"],
                       "us_per_launch_under_ncu": a["us"] / n,
                       "tflops_under_ncu": a["flops"] / (a["us"] * 1e-6) / 1e12 if a["us"] else None}
    tot_d = sum(a["dram_bytes"] for a in agg.values())
    tot_n = sum(a["launches"] for a in agg.values())
    res = {"source": f"{csv_path} x {shapes_path} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                     "one training step of the bench's shapes, cold-cache serialised launches)",
           "dram_bytes_per_launch": tot_d / tot_n, "launches": tot_n, "by_shape": by_shape}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1)[:3000])


def hbm(rep, out):
    """GB/s of every captured launch from its own dram bytes and duration (for the HBM-bound kernels)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    need = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"]
    with open(out, "w") as f:
        f.write(f"# HBM-bound kernels: achieved DRAM bandwidth per launch ({rep})\n\n")
        f.write("dram bytes and duration from the same ncu --set full capture (--clock-control none); peak = the measured "
                "copy bandwidth in MEASURED_PEAKS.json (6572 GB/s).\n\n")
        f.write("| kernel | grid | dram read MB | dram write MB | duration us | GB/s | frac of 6572 | tensor pipe % | sm throughput % |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for row in rd[2:]:
            try:
                rdb = _to_bytes(row[col[need[0]]], units[col[need[0]]])
                wrb = _to_bytes(row[col[need[1]]], units[col[need[1]]])
                us = _to_us(row[col[need[2]]], units[col[need[2]]])
            except (KeyError, ValueError):
                continue
            name = re.sub(r"\(.*", "", row[col["Kernel Name"]])[:70]
            tens = next((row[i] for h, i in col.items() if "pipe_tensor" in h and "pct" in h and row[i]), "")
            smt = row[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]] if "sm__throughput.avg.pct_of_peak_sustained_elapsed" in col else ""
            gbs = (rdb + wrb) / (us * 1e-6) / 1e9
            f.write(f"| {name} | {row[col['Grid Size']]} | {rdb / 1e6:.1f} | {wrb / 1e6:.1f} | {us:.1f} | {gbs:.0f} | "
                    f"{gbs / 6572.2:.3f} | {tens} | {smt} |\n")
    print(open(out).read()[:4000])


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "gemm_traffic":
        gemm_traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full, "hbm": hbm}[mode](sys.argv[2], sys.argv[3])
