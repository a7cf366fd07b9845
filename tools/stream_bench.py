#!/usr/bin/env python
"""End-to-end time of StripedSimulator.run() (pinned host arrays in and out) with the plain and the streamed schedule,
for several chunk counts:  python tools/stream_bench.py [--workload jacobi27] [--steps 100] [--chunks 8,16,32]
One JSON line per configuration; every streamed result is compared with the plain one (64-bit checksum of the host copy)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="jacobi27", choices=["jacobi27", "jacobi7"])
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--chunks", default="8,16,32")
    ap.add_argument("--dims", default="", help="x,y,z (default: the workload's 1024^3)")
    args = ap.parse_args()
    import torch
    from libgeodecomp_b200 import capi
    from libgeodecomp_b200.striping import StripedSimulator
    model_name, dims = bench.WORKLOADS[args.workload][:2]
    if args.dims:
        dims = tuple(int(v) for v in args.dims.split(","))
    depth = bench.TB_DEPTH.get(args.workload, 1)
    capi.set_tuning("jacobi.tb", depth)
    keep = []

    def alloc(shape, dtype):
        t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
        keep.append(t)
        return t.numpy()

    model, host = bench.synth_members(args.workload, dims, 0, dims[2], alloc)
    start = {n: a.copy() for n, a in host.items()}     # pageable copy of the input, to restart every run from it
    Init, Pull = bench.make_plugins(host, host, 2, 0)
    cells = float(np.prod(dims))
    want = None
    for chunks in [0] + [int(c) for c in args.chunks.split(",")]:
        for n, a in host.items():
            a[...] = start[n]
        sim = StripedSimulator(Init(dims, args.steps), model, device=torch.cuda.current_device(), stream_io=chunks > 0,
                               stream_depth=depth, stream_chunks=max(chunks, 1))
        sim.writers = [Pull("", 1 << 30)]
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        sim.run()
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        check = int(next(iter(host.values())).reshape(-1).view(np.int64).sum(dtype=np.int64))
        if want is None:
            want = check
        print(json.dumps({"workload": args.workload, "dims": list(dims), "steps": args.steps,
                          "schedule": "streamed, %d chunks" % chunks if chunks else "plain", "streamed_runs": sim.streamed_runs,
                          "ms_per_run": ms, "wall_ms": 1e3 * wall, "e2e_glups": 1e-9 * cells * args.steps / (1e-3 * ms),
                          "equals_plain": check == want}), flush=True)
        del sim


if __name__ == "__main__":
    main()
