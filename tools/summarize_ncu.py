"""Summarise ncu outputs into small text files under profiles/ (run in the build container, no GPU needed).

  python tools/summarize_ncu.py launches gpurun_out/r1a_launches.csv profiles/r1_launches_summary.md
  python tools/summarize_ncu.py full gpurun_out/r1a_prof_gemm.ncu-rep profiles/r1_gemm_full.md
"""
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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
