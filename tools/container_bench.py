"""The ContainerCell leg of bench.py on its own (one GPU): python tools/container_bench.py [--steps K] [--tile T] [--reps R]
prints the leg's JSON record. Used for the ncu captures of csrc/container.cu (profiles/)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--tile", type=int, default=128)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kernel", type=int, default=None, help='"container.kernel": 0 = windows / global gathers, 1 = tiles / staged values')
    args = ap.parse_args()
    import torch
    torch.cuda.set_device(0)
    print(json.dumps(bench.bench_container(args, torch, tile=args.tile, reps=args.reps, with_e2e=not args.no_e2e, kernel=args.kernel)), flush=True)


if __name__ == "__main__":
    main()
