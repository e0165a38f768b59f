"""BASELINE configs[1]: Jukebox front end only (see bench.jukebox_measure, which bench.py also folds into its JSON line
as `configs1_jukebox`).

    python tools/bench_jukebox.py [--precision tf32|bf16] [--batch 8] [--steps 3] [--depth 36]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--depth", type=int, default=36)
    a = ap.parse_args()
    print(json.dumps(bench.jukebox_measure(a.precision, a.batch, a.steps, a.warmup, a.depth)), flush=True)


if __name__ == "__main__":
    main()
