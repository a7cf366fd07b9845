#!/usr/bin/env python
"""Benchmark of the cell-update hot path (BASELINE.json metric: GLUPS, % of HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload jacobi27|jacobi7|lbm|gol|jacobi7_128]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
  python bench.py --impl reference ...      the reference's own CPU implementation (oracle/_ref)

A bench "step" is one sweep of the update over the whole (per-rank) grid. The headline workload is
BASELINE.json config 3 (Jacobi 27-point, 1024^3 f64 per GPU, slabs along z, halo exchange between
slab neighbours; weak scaling: the global grid is 1024 x 1024 x 1024N). The grids (2 x 8.9 GB) are
far larger than the 126 MB L2, so no explicit L2 flush is needed between timed iterations.

value  = lattice updates of ALL ranks / max-over-ranks device time, inputs resident in HBM.
e2e    = the same metric through the public API (StripedSimulator.run(): Initializer::grid from
         PINNED HOST memory -> K steps -> Writer pulling the final grid back to host memory), all
         host<->device copies inside the timed region; h2d/d2h bytes are per step (total / K).
         The large Jacobi workloads are timed a second time with stream_io (the same call; upload, sweeps and
         download pipelined chunk by chunk along z, striping.py::_run_streamed; on N > 1 ranks with ghost zones as
         wide as the run is long, so that no exchange is needed while the wavefront passes). e2e is the faster of
         the two schedules among those whose result was verified (element-wise against the other schedule, windows
         against the oracle); both are in the line (e2e.plain_schedule / e2e.streamed_schedule).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (model, per-rank dims (x, y, z), algorithmic bytes per update, dtype, ref model)
    "jacobi27": ("Jacobi27Cube", (1024, 1024, 1024), 16, "f64", "jacobi27cube"),
    "jacobi7": ("Jacobi7Cube", (1024, 1024, 1024), 16, "f64", "jacobi7cube"),
    "jacobi7_128": ("Jacobi7Cube", (128, 128, 128), 16, "f64", "jacobi7cube"),
    "lbm": ("LBMCellF", (512, 512, 512), 152, "f32", "lbm"),
    "gol": ("ConwayCube", (16384, 16384), 2, "u8", "conwaycube"),
}
DESCRIPTION = {
    "jacobi27": "Jacobi3D 27-point 1024^3 double per GPU, slab decomposition along z, halo exchange (BASELINE.json configs[2])",
    "jacobi7": "Jacobi3D 7-point 1024^3 double per GPU, slab decomposition along z",
    "jacobi7_128": "Jacobi3D 7-point 128^3 double (BASELINE.json configs[0], L2-resident)",
    "lbm": "LBM D3Q19 BGK lid-driven cavity 512^3 float SoA per GPU (BASELINE.json configs[3])",
    "gol": "Conway Game of Life 16384^2 char grid (BASELINE.json configs[1])",
}


# temporal blocking depth per workload (b200geo_set_tuning "jacobi.tb"): sweeps fused into one launch of
# the TMA-staged Jacobi kernel; with N > 1 the ghost zone must be that wide (one exchange per launch)
TB_DEPTH = {"jacobi27": 2, "jacobi7": 4, "jacobi7_128": 1, "lbm": 2}


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except (OSError, ValueError):
        return None


def measure_traffic(workload, depth):
    """DRAM bytes one launch of the dominant kernel moves, measured NOW on this box: a child process (tools/few_launches.py:
    the same grid, the same kernel through the same C ABI call) under `ncu --metrics dram__bytes_read.sum,
    dram__bytes_write.sum`, third launch. Only the byte counters are taken from the profiler, never a time."""
    kernel = {"jacobi27": "jacobi_tb", "jacobi7": "jacobi_tb", "lbm": "lbm_tb2" if depth > 1 else "lbm_kernel",
              "jacobi7_128": "jacobi"}.get(workload)
    if kernel is None:
        return None
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:" + kernel,
           "-s", "2", "-c", "1", "--csv", sys.executable, os.path.join(ROOT, "tools", "few_launches.py"), workload,
           ("lbm.tb=%d" if workload == "lbm" else "jacobi.tb=%d") % depth, "--sweeps", str(4 * max(1, depth))]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    except (OSError, subprocess.TimeoutExpired):
        return None
    read = written = None
    for line in res.stdout.splitlines():
        parts = [p.strip('"') for p in line.split('","')]
        if len(parts) < 3:
            continue
        try:
            value = float(parts[-1].replace(",", "").strip('"'))
        except ValueError:
            continue
        unit = parts[-2].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit)
        if scale is None:
            continue
        if "dram__bytes_read.sum" in line:
            read = value * scale
        elif "dram__bytes_write.sum" in line:
            written = value * scale
    if read is None or written is None:
        return None
    return {"dram_bytes_per_launch": read + written, "dram_bytes_read": read, "dram_bytes_written": written, "sweeps_per_launch": depth,
            "source": "measured in this run: ncu --metrics dram__bytes_{read,write}.sum on one launch of %s (child process)" % kernel}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def numa_bind(device_index):
    """Run this process on the CPU cores of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are
    allocated (first touch puts them on that node): host <-> device copies then do not cross the socket interconnect.
    Returns what was done (for the line's config); a box without that information is left alone."""
    info = {"bound": False}
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update({"bound": True, "cpus": len(cpus)})
    except (OSError, ValueError, AttributeError, RuntimeError) as e:
        info["error"] = type(e).__name__
    return info


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines, self.begin = index, None, [], 0

    def mark_begin(self):
        self.begin = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[self.begin:] or self.lines[-3:]
        for line in lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_members(workload, dims, z0, nz_total, alloc, share=False):
    """Synthetic input for one rank written straight into buffers from alloc(shape, dtype) (pinned
    host memory for the e2e leg), so that no second full-size copy ever exists on the host."""
    from libgeodecomp_b200 import models, synth
    name = WORKLOADS[workload][0]
    model = models.ALL[name]
    if name.startswith("Jacobi"):
        nx, ny, nz = dims
        tile = min(nz, 32)   # generate a 32-plane block and repeat it along z (generation speed)
        block = synth.jacobi_grid(nx, ny, tile, seed=42 + z0)
        out = alloc((nz, ny, nx), np.float64)
        for z in range(0, nz, tile):
            n = min(tile, nz - z)
            out[z:z + n] = block[:n]
        return model, {"temp": out}
    if name == "LBMCellF":
        nx, ny, nz = dims
        states = synth.lbm_states(nx, ny, nz, z0, nz_total)
        members, shared = {}, {}
        for m, (n, t) in enumerate(model.members):
            value = None if n == "state" else (1.0 if n in ("C", "density") else 0.0)
            if share and value in shared:      # read-only input: equal members may share one host array
                members[n] = shared[value]
                continue
            a = alloc((nz, ny, nx), t)
            a[...] = states if n == "state" else value
            members[n] = a
            if value is not None:
                shared[value] = a
        return model, members
    nx, ny = dims
    tile = min(ny, 1024)
    block = synth.gol_grid(nx, tile)
    out = alloc((ny, nx), np.uint8)
    for y in range(0, ny, tile):
        n = min(tile, ny - y)
        out[y:y + n] = block[:n]
    return model, {"alive": out}


def initial_window(workload, dims, world, lo, hi):
    """The synthetic input of synth_members on the global box [lo, hi) (array order: slowest axis first, global
    coordinates), regenerated without the full grid: {member: array}."""
    from libgeodecomp_b200 import models, synth
    name = WORKLOADS[workload][0]
    model = models.ALL[name]
    if name.startswith("Jacobi"):
        nx, ny, nz = dims
        tile = min(nz, 32)
        out = np.empty(tuple(h - l for l, h in zip(lo, hi)), dtype=np.float64)
        gz = np.arange(lo[0], hi[0])
        for r in sorted(set(gz // nz)):          # every rank repeats its own 32-plane block (seed 42 + z0)
            sel = np.nonzero(gz // nz == r)[0]
            out[sel] = synth.jacobi_box(nx, ny, tile, (0, lo[1], lo[2]), (0, hi[1], hi[2]), seed=42 + int(r) * nz,
                                        zmap=(gz[sel] % nz) % tile)
        return {"temp": out}
    if name == "LBMCellF":
        nx, ny, nz = dims
        shape = tuple(h - l for l, h in zip(lo, hi))
        out = {}
        for n, t in model.members:
            if n == "state":
                out[n] = synth.lbm_states_box(nx, ny, nz * world, lo, hi)
            else:
                out[n] = np.full(shape, 1.0 if n in ("C", "density") else 0.0, dtype=t)
        return out
    nx, ny = dims
    tile = min(ny, 1024)
    block = synth.gol_grid(nx, tile)
    gy = np.arange(lo[0], hi[0])
    return {"alive": np.ascontiguousarray(block[(gy % ny) % tile][:, lo[1]:hi[1]])}


def verify_windows(workload, dims, world, rank, host_out, steps):
    """The checker of the e2e leg: this rank's pulled result (host arrays, the rank's slab) against the oracle on
    WINDOWS of the global grid — at both slab faces (cells that depend on the neighbours' planes through `steps` halo
    exchanges) and inside. The update is local, so the oracle runs on the window plus a halo of `steps` cells wherever the
    window side is not a true domain boundary (tests/test_fullsize_gpu.py uses the same method). Bit-exact or False."""
    from oracle import oracle_py
    from libgeodecomp_b200 import models
    name = WORKLOADS[workload][0]
    model = models.ALL[name]
    nd = len(dims)
    ext = list(dims[::-1])                       # per-rank extents, slowest axis first
    gext = [ext[0] * world] + ext[1:]
    z0 = ext[0] * rank
    nz = ext[0]
    dz = min(4, nz)

    def clip(a, n, size):
        a = max(0, min(a, n - size))
        return a, min(n, a + size)

    inner = [clip(n // 2 - 7, n, 14) for n in ext[1:-1]] + [clip(ext[-1] // 3, ext[-1], 24)]
    corner = [clip(n, n, 12) for n in ext[1:-1]] + [clip(ext[-1], ext[-1], 20)]
    windows = [[(0, dz)] + inner, [(nz - dz, nz)] + corner, [clip(nz // 2 - 2, nz, 4)] + inner]
    checked = 0
    for win in windows:
        glo = [z0 + win[0][0]] + [a for a, _ in win[1:]]
        ghi = [z0 + win[0][1]] + [b for _, b in win[1:]]
        lo = [max(0, a - steps) for a in glo]
        hi = [min(n, b + steps) for b, n in zip(ghi, gext)]
        members = initial_window(workload, dims, world, lo, hi)
        inner_sl = tuple(slice(a - l, b - l) for a, b, l in zip(glo, ghi, lo))
        mine_sl = (slice(win[0][0], win[0][1]),) + tuple(slice(a, b) for a, b in win[1:])
        if name.startswith("Jacobi"):
            want = {"temp": oracle_py.jacobi(int(name[6:-4]), False, members["temp"], steps)}
        elif name == "LBMCellF":
            raw = np.stack([members[n].view(np.float32) for n, _ in model.members])
            res = oracle_py.lbm(raw, steps)
            want = {n: res[m].view(t) for m, (n, t) in enumerate(model.members)}
        else:
            want = {"alive": oracle_py.gol(False, members["alive"], steps)}
        for n, w in want.items():
            got = host_out[n][mine_sl]
            if not np.array_equal(got.view(np.uint8), np.ascontiguousarray(w[inner_sl]).view(np.uint8)):
                return {"ok": False, "rank": rank, "window": [glo, ghi], "member": n}
        checked += int(np.prod([b - a for a, b in zip(glo, ghi)]))
    return {"ok": True, "rank": rank, "windows": len(windows), "cells": checked}


def mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1048576.0
    except OSError:
        pass
    return 0.0


def run_cpu_reference(workload, steps, warmup, threads=None, full=True, full_steps=20):
    """Time the reference's own OpenMPSimulator (oracle/_ref) on the workload. The Jacobi 1024^3 configurations run at
    FULL size when the host has the memory (two 1032^3 f64 grids = 17.6 GB; the synthetic input repeats every 32 planes,
    so only that block is handed over) — SURVEY.md 8(d) C3 — with the step count capped so that the run ends within a
    few minutes; otherwise, and for the other workloads, on a bounded sample. Returns (glups, dict)."""
    from oracle import oracle_py
    from libgeodecomp_b200 import synth
    model_name, dims, _, _, ref_model = WORKLOADS[workload]
    cores = threads or os.cpu_count()
    tile, same_config = 0, False
    if workload in ("jacobi27", "jacobi7"):
        if full and oracle_py.have_ref(ref_model) and mem_available_gb() >= 30.0:
            sdims, tile, same_config = dims, 32, True
            steps, warmup = min(steps, full_steps), 0
            sample = "the full 1024^3 grid (same config as the device arm)"
            raw = synth.jacobi_grid(dims[0], dims[1], tile)
        else:
            sdims, sample = (1024, 1024, 32), "1024x1024x32 slab of the 1024^3 grid"
            raw = synth.jacobi_grid(*sdims)
    elif workload == "jacobi7_128":
        sdims, sample, same_config = (128, 128, 128), "full 128^3 grid", True
        raw = synth.jacobi_grid(*sdims)
    elif workload == "lbm":
        sdims, sample = (512, 512, 8), "512x512x8 slab of the 512^3 grid"
        raw = synth.lbm_grid(*sdims)
    else:
        sdims, sample = (16384, 256, 1), "16384x256 strip of the 16384^2 grid"
        raw = synth.gol_grid(sdims[0], sdims[1])
    if oracle_py.have_ref(ref_model):
        if warmup:
            oracle_py.run_ref(ref_model, raw, sdims, warmup, omp=True, threads=cores, want_output=False, tile=tile)
        _, st = oracle_py.run_ref(ref_model, raw, sdims, steps, omp=True, threads=cores, want_output=False, tile=tile)
        glups = st["glups_compute"]
        info = {"kind": "reference", "cores": cores, "simulator": st["simulator"],
                "sample": "%s x %d steps, reference OpenMPSimulator built from /root/reference, TimeCompute interval" % (sample, steps),
                "seconds": st["time_compute_s"], "steps": steps, "same_config": same_config}
    else:
        fn = {"jacobi27": lambda: oracle_py.jacobi(27, False, raw, steps),
              "jacobi7": lambda: oracle_py.jacobi(7, False, raw, steps),
              "jacobi7_128": lambda: oracle_py.jacobi(7, False, raw, steps),
              "lbm": lambda: oracle_py.lbm(raw, steps),
              "gol": lambda: oracle_py.gol(False, raw, steps)}[workload]
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        glups = 1e-9 * steps * float(np.prod(sdims)) / dt
        info = {"kind": "port", "cores": cores, "sample": "%s x %d steps, C restatement (oracle/oracle.c, OpenMP)" % (sample, steps),
                "seconds": dt, "steps": steps, "same_config": same_config}
    info.update({"value": glups, "unit": "GLUPS"})
    return glups, info


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    glups, info = run_cpu_reference(args.workload, args.steps, args.warmup)
    model_name, dims, _, dtype, _ = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "GLUPS (giga lattice updates/s)", "value": glups, "unit": "GLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * info["seconds"] / max(1, info["steps"]), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": DESCRIPTION[args.workload], "cell": model_name,
                   "dims_per_gpu": list(dims), "same_config": info["same_config"],
                   "note": "CPU reference on the host cores of this box; see cpu_baseline.sample for what was timed "
                           "(%d of the %d steps asked for)" % (info["steps"], args.steps)},
        "cpu_baseline": info,
        "e2e": {"value": glups, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


def bench_device(workload, args, rank, world, dist, torch, with_e2e=True, with_clocks=True):
    """Returns a dict with value / roofline / e2e for one workload."""
    from libgeodecomp_b200 import capi
    from libgeodecomp_b200.striping import StripedSimulator

    model_name, dims, alg_bytes, dtype, _ = WORKLOADS[workload]
    last = len(dims) - 1
    gdims = list(dims)
    gdims[last] = dims[last] * world
    z0 = dims[last] * rank
    keep, padded = [], []
    # slabs whose e2e leg will be streamed get their ghost zones (as wide as the run is long) in the SAME pinned
    # allocation, around the slab: one page-locked buffer per member and rank, not two
    pad_lo, pad_hi = stream_pads(workload, args, rank, world, dims) if with_e2e else (0, 0)

    def alloc(shape, dtype):
        t = torch.empty((shape[0] + pad_lo + pad_hi,) + tuple(shape[1:]), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=with_e2e)
        keep.append(t)
        padded.append(t.numpy())
        return t.numpy()[pad_lo:pad_lo + shape[0]]

    model, members = synth_members(workload, dims, z0, gdims[last], alloc, share=not with_e2e)
    ext = dict(zip(members.keys(), padded)) if with_e2e else None
    cells_rank = float(np.prod(dims))
    cells_all = cells_rank * world
    K, W = args.steps, args.warmup
    # the e2e leg reads its input from these (pinned) host buffers and writes the result back into them
    pinned = members
    host_out = members

    Init, PullWriter = make_plugins(pinned, host_out, last, z0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    depth = TB_DEPTH.get(workload, 1)
    capi.set_tuning("jacobi.tb", depth)
    ghost = args.ghost if args.ghost else depth
    sim = StripedSimulator(Init(gdims, K), model, rank=rank, world=world, ghost_width=ghost, device=torch.cuda.current_device(), dist=dist)
    torch.cuda.synchronize()

    # ---- device-resident throughput
    sampler = ClockSampler(torch.cuda.current_device())
    if with_clocks and rank == 0:
        sampler.start()
    sim.advance(W)
    barrier()
    sampler.mark_begin()
    launches0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sim.advance(K)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = capi.launch_count() - launches0
    clocks = sampler.stop() if (with_clocks and rank == 0) else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = 1e-9 * cells_all * K / (1e-3 * ms)

    # ---- dominant kernel alone (roofline): K sweeps without any exchange on a 1-rank-style step
    peak, peak_src = peaks()
    launches_roof0 = capi.launch_count()
    ev0.record()
    n_roof = 0
    while n_roof < K:
        n = K - n_roof
        if world > 1:   # no exchange here: the ghost planes are only declared valid (timing only)
            n = min(n, sim.ghost_width)
            for side in (0, 1):
                if sim.grid.modes[last][side] == capi.GHOST_PEER:
                    sim.grid.dev.halo_mark_valid(side, sim.ghost_width)
        sim.grid.dev.step(model.kernel, n_steps=n, params=model.step_params(n_roof + n == K))
        n_roof += n
    ev1.record()
    torch.cuda.synchronize()
    n_launch = capi.launch_count() - launches_roof0
    kms = ev0.elapsed_time(ev1) / n_launch               # mean duration of one launch
    per_launch = n_roof / n_launch                       # sweeps per launch (temporal blocking depth)
    achieved = alg_bytes * cells_rank * per_launch / (1e-3 * kms) / 1e9
    traffic = None
    if with_clocks and rank == 0 and world == 1 and not args.no_cpu:      # the headline workload of a default run
        traffic = measure_traffic(workload, int(round(per_launch)))
    traffic = traffic or ncu_traffic(workload) or {}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes_per_launch"), "traffic_source": traffic.get("source"),
                "kernel_ms": kms, "sweeps_per_launch": per_launch, "updates_per_launch": cells_rank * per_launch,
                "algorithmic_bytes_per_update": alg_bytes, "peak_source": peak_src, "kernel": model_name}
    if roofline["traffic"] and traffic.get("sweeps_per_launch") and abs(traffic["sweeps_per_launch"] - per_launch) < 1e-9:
        # what the DRAM actually moved per launch / launch time: the physical HBM fraction (<= 1);
        # with temporal blocking the ALGORITHMIC rate above may exceed the copy peak, this one cannot
        roofline["dram_frac"] = roofline["traffic"] / (1e-3 * kms) / 1e9 / peak

    out = {"workload": workload, "value": value, "ms_per_step": ms / K, "roofline": roofline, "dtype": dtype,
           "gpu_launches": launches, "clocks": clocks, "model": model_name, "dims_per_gpu": list(dims),
           "global_dims": gdims, "halo_bytes_per_exchange": sim.halo.bytes_per_exchange,
           "ghost_width": sim.ghost_width}

    # ---- end to end through the public API with host buffers
    if with_e2e:
        out.update(e2e_legs(workload, args, rank, world, dist, torch, sim, model, members, ext, (pad_lo, pad_hi), dims, gdims, z0, depth, barrier))
    del sim
    return out


def make_plugins(pinned, host_out, last, z0):
    """The Initializer and the Writer of the e2e leg: host arrays (pinned) in, host arrays out, through the plugin API."""
    from libgeodecomp_b200.simulator import ParallelWriter, SimpleInitializer

    class Init(SimpleInitializer):
        """initialises the cells inside target.boundingBox() — the rank's slab, or one window of it when run() streams"""
        def grid(self, target):
            (o, d) = target.boundingBox()
            a0 = o[last] - z0
            for n, a in pinned.items():
                target.loadMember(n, a[a0:a0 + d[last]], origin=o)

    class PullWriter(ParallelWriter):
        """pulls the rank's final grid into pinned host memory at WRITER_ALL_DONE, region by region"""
        def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank_, lastCall):
            if event == 2:
                (o, d) = validRegion
                a0 = o[last] - z0
                for n, a in host_out.items():
                    grid.saveMember(n, origin=o, dims=d, out=a[a0:a0 + d[last]])

    return Init, PullWriter


def stream_wanted(workload, args, world, dims):
    """does the e2e leg of this workload get a streamed run? Large Jacobi grids on a Cube; on slabs only while ghost zones
    as wide as the run add at most 25 % to the slab"""
    from libgeodecomp_b200 import models
    model = models.ALL[WORKLOADS[workload][0]]
    nz = dims[-1]
    grid_bytes = float(np.prod(dims)) * sum(t.itemsize for _, t in model.members)
    if not getattr(model, "fuses_sweeps", False) or model.wraps or grid_bytes < (1 << 30) or args.no_stream:
        return False
    return world == 1 or 8 * args.steps * model.nano_steps <= nz


def stream_pads(workload, args, rank, world, dims):
    """ghost planes (low, high) a rank's pinned input arrays carry for the streamed leg"""
    from libgeodecomp_b200 import models
    if world == 1 or not stream_wanted(workload, args, world, dims):
        return 0, 0
    G = args.steps * models.ALL[WORKLOADS[workload][0]].nano_steps
    return (G if rank > 0 else 0), (G if rank < world - 1 else 0)


def all_ranks(flag, world, dist, torch):
    """[flag of rank 0, flag of rank 1, ...] on every rank"""
    t = torch.tensor([1 if flag else 0], dtype=torch.int64, device="cuda")
    if world == 1:
        return [bool(flag)]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [bool(int(x.item())) for x in out]


def max_over_ranks(ms, world, dist, torch):
    t = torch.tensor([max(ms, 0.0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def host_link(world, dist, torch, barrier, nbytes=1 << 30):
    """What the links between host memory and the GPUs carry on THIS box with all ranks copying at once (page-locked
    memory, CUDA events, max over ranks): the ceiling of every e2e number. One GiB per direction and rank."""
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        fn()                      # warm-up
        barrier()
        ev0.record()
        fn()
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1), world, dist, torch)

    def duplex():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            host2.copy_(dev2, non_blocking=True)
        dev.copy_(host, non_blocking=True)
        torch.cuda.current_stream().wait_stream(side)

    up = timed(lambda: dev.copy_(host, non_blocking=True))
    down = timed(lambda: host.copy_(dev, non_blocking=True))
    both = timed(duplex)
    gb = nbytes / 1e9
    return {"ranks": world, "h2d_gbs_per_rank": gb / (1e-3 * up), "d2h_gbs_per_rank": gb / (1e-3 * down),
            "duplex_gbs_per_rank": 2 * gb / (1e-3 * both), "aggregate_duplex_gbs": 2 * gb * world / (1e-3 * both),
            "what": "1 GiB per direction and rank, page-locked, all ranks at once, max over ranks"}


def e2e_legs(workload, args, rank, world, dist, torch, sim, model, members, ext, pads, dims, gdims, z0, depth, barrier):
    """The e2e legs of one workload: StripedSimulator.run() from pinned host arrays back into pinned host arrays,
    (1) with the plain schedule (upload, sweeps with halo exchanges, download), (2) — large Jacobi grids — with the
    streamed schedule. Every result is checked: windows against the oracle on every rank, and the two schedules'
    final grids against each other element by element. `e2e` = the faster verified schedule."""
    from libgeodecomp_b200 import capi
    from libgeodecomp_b200.striping import StripedSimulator
    K = args.steps
    last = len(dims) - 1
    nz = dims[last]
    cells_all = float(np.prod(dims)) * world
    grid_bytes = sum(int(x.nbytes) for x in members.values())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Init, PullWriter = make_plugins(members, members, last, z0)
    res = {}

    def timed_run(simulator):
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        simulator.run()
        ev1.record()
        wall = time.perf_counter() - t0      # run() has returned: the host buffers must be complete by now
        return wall

    def checked(host_out, steps):
        if args.no_verify:
            return None
        ver = verify_windows(workload, dims, world, rank, host_out, steps)
        per_rank = all_ranks(ver["ok"], world, dist, torch)
        v = {"ok": all(per_rank), "per_rank": per_rank, "steps": steps, "cells_per_rank": ver.get("cells", 0)}
        if not ver["ok"]:
            v["first_mismatch"] = ver
        return v

    def entry(e_ms, wall, what):
        return {"value": 1e-9 * cells_all * K / (1e-3 * e_ms), "unit": "GLUPS", "h2d_bytes_per_step": grid_bytes * world / K,
                "d2h_bytes_per_step": grid_bytes * world / K, "ms_per_run": e_ms, "wall_ms_rank0": 1e3 * wall, "what": what}

    # ---- (1) the plain schedule
    sim.writers = [PullWriter("", 1 << 30)]
    sim.initializer = Init(gdims, K)
    wall = timed_run(sim)
    barrier()
    plain = entry(max_over_ranks(ev0.elapsed_time(ev1), world, dist, torch), wall,
                  "StripedSimulator.run(): Initializer::grid from pinned host memory -> %d steps -> Writer pulls the final "
                  "grid to pinned host memory" % K)
    plain["schedule"] = "plain: upload, sweeps (halo exchange every %d), download one after the other" % sim.ghost_width
    vsteps = K
    if K > 40 and not args.no_verify:
        # the oracle's halo grows with the step count: beyond 40 steps the same call is repeated (untimed) with 16 steps
        # on the regenerated input and THAT result is checked
        vsteps = 16
        spare = iter(list(members.values()))
        synth_members(workload, dims, z0, gdims[last], lambda shape, dtype: next(spare), share=False)
        sim.initializer = Init(gdims, vsteps)
        sim.run()
        barrier()
    try:
        link = host_link(world, dist, torch, barrier)
        # the time the run's host traffic alone takes at that rate: one way after the other (plain) / both ways at once (streamed)
        link["plain_floor_ms"] = 1e3 * (grid_bytes / 1e9) * (1 / link["h2d_gbs_per_rank"] + 1 / link["d2h_gbs_per_rank"])
        link["streamed_floor_ms"] = 1e3 * 2 * (grid_bytes / 1e9) / link["duplex_gbs_per_rank"]
        res["host_link"] = link
    except Exception as e:  # noqa: BLE001 - a diagnostic
        res["host_link"] = {"error": repr(e)}
    ver = checked(members, vsteps)
    if ver is not None:
        ver["what"] = ("the timed e2e run's pulled result" if vsteps == K else "an untimed repeat of the e2e call with %d steps" % vsteps) + \
                      ", bit for bit against the oracle on 3 windows per rank (both slab faces, interior)"
        plain["verified"] = ver["ok"]
        res["verified"] = ver
    res["e2e"] = plain

    # ---- (2) the streamed schedule
    G = K * model.nano_steps if world > 1 else 0     # ghost zones as wide as the run is long: no exchange while streaming
    pad_lo, pad_hi = pads
    wanted = stream_wanted(workload, args, world, dims)
    if not wanted and world > 1 and WORKLOADS[workload][0].startswith("Jacobi") and not args.no_stream and grid_bytes >= (1 << 30):
        res["e2e"]["streamed_schedule"] = {"skipped": "%d sweeps: ghost zones of that width would add more than 25 %% to a slab of %d planes" % (G, nz)}
    if not wanted:
        return res
    try:
        name = model.members[0][0]
        own = members[name]
        own_view = members      # views into `ext` (bench_device allocated slab and ghost zones in one piece)

        def refill():
            spare = iter(list(own_view.values()))
            synth_members(workload, dims, z0, gdims[last], lambda shape, dtype: next(spare), share=False)
            ext_lo = [z0 - pad_lo] + [0] * last
            for lo, hi in (([z0 - pad_lo], [z0]), ([z0 + nz], [z0 + nz + pad_hi])):
                if hi[0] > lo[0]:
                    glo, ghi = lo + [0] * last, hi + list(dims[::-1][1:])
                    for n, a in initial_window(workload, dims, world, glo, ghi).items():
                        ext[n][lo[0] - ext_lo[0]:hi[0] - ext_lo[0]] = a

        SInit, SPull = make_plugins(ext, ext, last, z0 - pad_lo)
        sim2 = StripedSimulator(SInit(gdims, min(K, 16)), model, rank=rank, world=world, ghost_width=(G if world > 1 else depth),
                                device=torch.cuda.current_device(), dist=dist, stream_io=True, stream_depth=depth)
        sim2.writers = [SPull("", 1 << 30)]
        # warm-up = the oracle check of the schedule: a streamed run of <= 16 steps, windows against the oracle
        refill()
        sim2.run()
        barrier()
        oracle = checked(own_view, min(K, 16))
        streamed_everywhere = all(all_ranks(sim2.streamed_runs == 1, world, dist, torch))
        # the plain schedule's K-step result from the same input, on the device
        want = torch.empty(own.shape, dtype=getattr(torch, own.dtype.name), device="cuda")
        if vsteps != K:
            spare = iter(list(members.values()))
            synth_members(workload, dims, z0, gdims[last], lambda shape, dtype: next(spare), share=False)
            sim.initializer = Init(gdims, K)
            sim.writers = []
            sim.run()
        sim.grid.saveMember(name, out=want, location=capi.CUDA_DEVICE)
        torch.cuda.synchronize()
        sample = sorted(set(list(range(0, nz, 64)) + [nz - 2, nz - 1]))
        want_sample = want[sample].cpu().numpy()
        # the timed streamed run
        refill()
        sim2.initializer = SInit(gdims, K)
        launches0 = capi.launch_count()
        wall = timed_run(sim2)
        # read by the CPU right after run() returned, with no synchronisation of our own (the last planes pulled are the
        # ones that would still be in flight): the ParallelWriter contract
        host_now = bool(np.array_equal(own_view[name][sample].view(np.uint8), want_sample.view(np.uint8)))
        barrier()
        launches = capi.launch_count() - launches0
        s_ms = max_over_ranks(ev0.elapsed_time(ev1), world, dist, torch)
        got = torch.empty_like(want)
        sim2.grid.saveMember(name, out=got, location=capi.CUDA_DEVICE)
        torch.cuda.synchronize()
        device_same = bool(torch.equal(got.view(torch.uint8), want.view(torch.uint8)))
        host_same = host_now
        step = max(1, nz // 16)
        for a in range(0, nz, step):
            got[a:a + step].copy_(torch.from_numpy(own_view[name][a:a + step]))
            host_same = host_same and bool(torch.equal(got[a:a + step].view(torch.uint8), want[a:a + step].view(torch.uint8)))
        del got, want
        final = checked(own_view, K) if K <= 40 else None
        ok = streamed_everywhere and all(all_ranks(device_same and host_same, world, dist, torch)) and \
            (oracle is None or oracle["ok"]) and (final is None or final["ok"])
        levels, chunk = sim2._stream_plan() or ([0], 0)
        extra = (pad_lo + pad_hi) * int(np.prod(dims[:last])) * sum(a.dtype.itemsize for a in ext.values())
        streamed = entry(s_ms, wall, "the same call with stream_io: Initializer -> sweeps -> ParallelWriters pipelined chunk by chunk")
        streamed.update({
            "h2d_bytes_per_step": (grid_bytes * world + (extra * world if world > 1 else 0)) / K,
            "schedule": "streamed: z-chunks of %d planes, %d levels of <= %d fused sweeps, time-skewed; upload / sweeps / download on "
                        "three streams%s" % (chunk, len(levels), max(levels),
                                             "" if world == 1 else "; ghost zones %d planes wide from the Initializer, no halo exchange" % G),
            "verified": ok, "gpu_launches": launches,
            "how": "final grid equal ELEMENT BY ELEMENT to the plain schedule's on the same input on every rank (device %s; host copy, "
                   "first read right after run() returned %s); a streamed run of %d steps bit-exact vs the oracle on 3 windows per "
                   "rank (%s)%s" % (device_same, host_same, min(K, 16), None if oracle is None else oracle["ok"],
                                    "" if final is None else "; the timed run's result likewise (%s)" % final["ok"])})
        del sim2
        if ok and streamed["value"] > plain["value"]:
            res["e2e"] = dict(streamed, plain_schedule={k: plain[k] for k in ("value", "ms_per_run", "schedule")})
        else:
            res["e2e"]["streamed_schedule"] = {k: streamed[k] for k in ("value", "ms_per_run", "schedule", "verified", "how")}
    except Exception as e:  # noqa: BLE001 - whatever goes wrong in the second leg, the plain number stands
        res["e2e"]["streamed_schedule"] = {"verified": False, "error": "%s: %s" % (type(e).__name__, e)}
    return res


def gpu_reference(workload, args, torch):
    """The reference's OWN GPU path on this B200 beside ours: LibGeoDecomp's CUDASimulator<CELL> (cudasimulator.h, AoS
    Jacobi cell with FixedCoord access) recompiled for sm_100a (oracle/_ref/lgd_ref_cuda_jacobi, built by oracle/Makefile
    from /root/reference) at 512^3 — its host-side DisplacedGrid initialisation makes 1024^3 impractical — and
    libb200geo.so on a grid of the same size, both timed with CUDA events around K steps of resident data."""
    from libgeodecomp_b200 import capi, models
    from libgeodecomp_b200.simulator import B200Grid
    kind = {"jacobi27": "27", "jacobi7": "7"}.get(workload)
    exe = os.path.join(ROOT, "oracle", "_ref", "lgd_ref_cuda_jacobi")
    if kind is None or not os.access(exe, os.X_OK):
        return None
    n, K = 512, max(10, min(args.steps, 50))
    res = subprocess.run([exe, str(n), str(K), kind], capture_output=True, text=True, timeout=300)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    if res.returncode != 0 or not lines:
        return {"error": "lgd_ref_cuda_jacobi exited with %d" % res.returncode}
    ref = json.loads(lines[-1])
    model = models.ALL[WORKLOADS[workload][0]]
    capi.set_tuning("jacobi.tb", TB_DEPTH.get(workload, 0))   # the secondary workloads set their own depth
    grid = B200Grid(model, (n, n, n))
    grid.loadMember("temp", torch.rand((n, n, n), dtype=torch.float64, device="cuda"), location=capi.CUDA_DEVICE)
    grid.dev.step(model.kernel, n_steps=4)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    grid.dev.step(model.kernel, n_steps=K)
    ev1.record()
    torch.cuda.synchronize()
    ours = 1e-9 * n ** 3 * K / (1e-3 * ev0.elapsed_time(ev1))
    del grid
    return {"impl": ref["impl"], "cell": ref["cell"], "dims": ref["dims"], "steps": K, "value": ref["glups"], "unit": "GLUPS",
            "ms_per_step": ref["ms_per_step"], "b200geo_same_dims": ours, "ratio": ours / ref["glups"] if ref["glups"] else None,
            "cuda_status": ref.get("cuda")}


def e2e_cpp(args):
    """The same e2e measurement through the C++ facade, the reference's host language: tests/facade/_bin/e2e_bench (a
    LibGeoDecomp program on B200Simulator: Initializer from / Writer into page-locked host memory, run() timed by the
    program itself) — handing the engine whole boxes with a serial Writer (plain schedule), with a ParallelWriter (the C++
    streamed schedule, b200streamedrun.h), and going row by row as BOVOutput::writeGrid does. `value` is the best of the
    box-wise schedules whose checksums agree."""
    exe = os.path.join(ROOT, "tests", "facade", "_bin", "e2e_bench")
    if args.workload != "jacobi27" or not os.access(exe, os.X_OK):
        return None
    out = {}
    for mode in ("box", "stream", "rows"):
        res = subprocess.run([exe, "1024", str(args.steps), "1", mode], capture_output=True, text=True, timeout=600)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            out[mode] = {"error": "e2e_bench exited with %d: %s" % (res.returncode, (res.stdout + res.stderr)[-200:])}
            continue
        out[mode] = json.loads(lines[0])
    if "value" in out.get("box", {}):
        out.update({"value": out["box"]["value"], "unit": "GLUPS", "schedule": "plain"})
        if "checksum" in out.get("rows", {}):
            out["rows_equal_box"] = out["rows"]["checksum"] == out["box"]["checksum"]
        if "checksum" in out.get("stream", {}):
            out["stream_equal_box"] = out["stream"]["checksum"] == out["box"]["checksum"]
            if out["stream_equal_box"] and out["stream"]["value"] > out["value"]:
                out.update({"value": out["stream"]["value"], "schedule": "streamed"})
    return out


def bench_nbody(args, rank, world, dist, torch, containers=108, with_e2e=True):
    """BASELINE.json configs[4]: short-range n-body in BoxCell containers, ~16.6 M particles per GPU
    (255^3 lattice sites in 108^3 containers of edge 2.5 = cutoff), slabs of containers along z.
    Metric: particle updates/s and candidate pair evaluations/s; bound = FP32 pipe."""
    from libgeodecomp_b200 import capi, models, synth
    from libgeodecomp_b200.simulator import SimpleInitializer, Writer
    from libgeodecomp_b200.striping import StripedSimulator

    n = containers
    model = models.NBodyF
    K, W = max(4, args.steps // 5), 2
    cap = model.capacity
    c, p = synth.nbody_cells(n, n, n, cap=cap, z0=rank * n, dtype=np.float32)
    pc = torch.empty(c.shape, dtype=torch.int32, pin_memory=True)
    pp = torch.empty(p.shape, dtype=torch.float32, pin_memory=True)
    pc.numpy()[...] = c
    pp.numpy()[...] = p
    hc, hp = pc.numpy(), pp.numpy()
    particles = int(c.sum())
    # candidate pairs of one sweep: every particle meets every particle of its 27 containers
    cc = np.pad(c.astype(np.int64), 1)
    hood = sum(cc[dz:dz + n, dy:dy + n, dx:dx + n] for dz in range(3) for dy in range(3) for dx in range(3))
    pairs = int((c.astype(np.int64) * hood).sum())
    del c, p

    class Init(SimpleInitializer):
        def grid(self, target):
            o, d = target.boundingBox()
            target.loadCells(hc, hp, origin=o)

    class Pull(Writer):
        def stepFinished(self, grid, step, event):
            if event == 2:
                grid.saveCells(out=(hc, hp))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sim = StripedSimulator(Init((n, n, n * world), K), model, rank=rank, world=world, ghost_width=1,
                           device=torch.cuda.current_device(), dist=dist)
    sim.advance(W)
    barrier()
    launches0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sim.advance(K)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    launches = capi.launch_count() - launches0
    sim.grid.dev.check()
    # FP32 pipe: 8 flop per candidate (3 sub, 3 mul, 2 add; the reference arithmetic has no FMA) and
    # ~20 more for the ~15 % inside the cutoff; peak = 148 SMs x 128 lanes x 2 flop x max SM clock
    flop = pairs * (8 + 0.155 * 20)
    peak = 148 * 128 * 2 * 1.965e9 / 1e12
    out = {"workload": "nbody", "metric": "G particle updates/s", "value": 1e-9 * particles * world / (1e-3 * ms), "unit": "Gparticles/s",
           "ms_per_step": ms, "particles_per_gpu": particles, "containers_per_gpu": [n, n, n], "capacity": cap,
           "candidate_pairs_per_s": pairs * world / (1e-3 * ms), "dtype": "f32", "gpu_launches": launches, "steps": K,
           "roofline": {"bound": "fp32-pipe", "achieved": flop / (1e-3 * ms) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": flop / (1e-3 * ms) / 1e12 / peak, "traffic": (ncu_traffic("nbody") or {}).get("dram_bytes_per_launch"),
                        "note": "algorithmic flop (8 per candidate pair + 20 per interacting pair, no FMA by the parity rule) "
                                "against the nominal FMA peak; pipe utilisation from ncu in profiles/"}}
    if with_e2e:
        sim.writers = [Pull("", 1 << 30)]
        barrier()
        ev0.record()
        sim.run()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
        nbytes = hc.nbytes + hp.nbytes
        out["e2e"] = {"value": 1e-9 * particles * world * K / (1e-3 * e_ms), "unit": "Gparticles/s",
                      "h2d_bytes_per_step": nbytes * world / K, "d2h_bytes_per_step": nbytes * world / K, "ms_per_run": e_ms,
                      "what": "StripedSimulator.run(): containers from pinned host memory -> %d steps -> pulled back" % K}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:   # the reference's own OpenMPSimulator over BoxCell containers on a bounded sample
            from oracle import oracle_py
            sc, sp = synth.nbody_cells(24, 24, 24, dtype=np.float32)
            _, st = oracle_py.run_ref_nbody(sc, sp, 3, omp=True, threads=os.cpu_count(), want_output=False)
            out["cpu_baseline"] = {"value": st["gpups_compute"], "unit": "Gparticles/s", "cores": os.cpu_count(), "kind": "reference",
                                   "sample": "24^3 containers (%d particles) x 3 steps, reference OpenMPSimulator over "
                                             "BoxCell<FixedArray<LJParticle<float>, 32>>, TimeCompute interval" % int(sc.sum())}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (e,)}
    del sim
    return out


def replicate_mesh(torch, box, tile, reps, dev):
    """A `tile`^2 torus mesh of synth.container_cells() replicated reps x reps times into one (tile * reps)^2 torus,
    built with torch on `dev`: element ids follow the global container index (1 + container * cap + slot), neighbour
    references keep their offset to the neighbour container and their slot. Returns (dict of the six arrays, the
    function that replicates a per-tile array)."""
    cap, maxnb = box["nb_ids"].shape[-2:]
    n = tile * reps
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in box.items()}
    # the tile's neighbour references as (offset to the neighbour container, slot), re-based on every copy of the tile
    ly, lx = torch.meshgrid(torch.arange(tile, device=dev), torch.arange(tile, device=dev), indexing="ij")
    ref = t["nb_ids"].long() - 1
    tc, slot = ref // cap, ref % cap
    dx = (tc % tile - lx[:, :, None, None] + 1) % tile - 1
    dy = (tc // tile - ly[:, :, None, None] + 1) % tile - 1
    live = torch.arange(maxnb, device=dev)[None, None, None, :] < t["nb_counts"][..., None]
    gy, gx = torch.meshgrid(torch.arange(n, device=dev), torch.arange(n, device=dev), indexing="ij")

    def rep(a):
        return a.repeat((reps, reps) + (1,) * (a.dim() - 2))

    g = {k: rep(t[k]).contiguous() for k in ("counts", "values", "influx", "nb_counts")}
    g["ids"] = (1 + (gy * n + gx)[:, :, None] * cap + torch.arange(cap, device=dev)[None, None, :]).int()
    g["ids"][torch.arange(cap, device=dev)[None, None, :] >= g["counts"][..., None]] = 0
    tx = (gx[:, :, None, None] + rep(dx)) % n
    ty = (gy[:, :, None, None] + rep(dy)) % n
    g["nb_ids"] = torch.where(rep(live), 1 + (ty * n + tx) * cap + rep(slot), 0).int().contiguous()
    return g, rep


def bench_container(args, torch, tile=128, reps=8, with_e2e=True, kernel=None):
    """ContainerCell grid of ID-keyed mesh elements (the cell of the reference's Voronoi example; SURVEY 8(a) a10's
    ID-keyed variant): a 2-D torus of (tile * reps)^2 containers, capacity 16, up to 20 neighbour ids per element.
    The mesh is a `tile`^2 torus replicated reps x reps times (built on the device), so the full-size result must
    equal the small torus's result tile by tile, and THAT is checked against the oracle: a size-independent check of
    every temperature. Metric: element updates/s; bound = HBM (the link table is read once per sweep)."""
    from libgeodecomp_b200 import capi, models, synth
    from libgeodecomp_b200.containergrid import ContainerGrid

    model = models.Container2Torus
    cap, maxnb = model.capacity, model.max_neighbors
    K, W = max(10, args.steps // 2), 3
    if kernel is not None:
        capi.set_tuning("container.kernel", kernel)
    box, _ = synth.container_cells(tile, tile, 1, n_dims=2, torus=True, cap=cap, maxnb=maxnb, seed=11)
    n = tile * reps
    dev = "cuda"
    g, rep = replicate_mesh(torch, box, tile, reps, dev)
    elements = int(g["counts"].sum().item())
    links = int(box["nb_counts"][np.arange(cap) < box["counts"][..., None]].sum()) * reps * reps

    grid = ContainerGrid(model, (n, n))
    grid.dev.load(g, location=capi.CUDA_DEVICE)
    grid.dev.step(capi.KERNEL_CONTAINER, W)
    torch.cuda.synchronize()
    launches0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    grid.dev.step(capi.KERNEL_CONTAINER, K)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / K
    launches = capi.launch_count() - launches0
    stats = grid.dev.stats()
    assert stats["cargo"] == elements and stats["links"] == links, (stats, elements, links)
    # every temperature: the replicated mesh must evolve like the tile torus, which the oracle computes
    out = torch.empty((n, n, cap), dtype=torch.float64, device=dev)
    grid.dev.save({"values": out}, location=capi.CUDA_DEVICE)
    verified = None
    if not args.no_verify:
        from oracle import oracle_py
        want = torch.from_numpy(oracle_py.container(box, W + K, n_dims=2, torus=True)).to(dev)
        verified = bool(torch.equal(out.view(torch.int64), rep(want).contiguous().view(torch.int64)))
    # algorithmic bytes of one sweep: the link (4 B each), and per element its influx (8), neighbour count (4), the
    # write (8) and one compulsory read of its own old temperature (8): the gathers re-read what other elements fetch
    alg = 4.0 * links + 28.0 * elements
    peak, peak_src = peaks()
    res = {"workload": "container", "metric": "G element updates/s", "value": 1e-9 * elements / (1e-3 * ms), "unit": "Gelements/s",
           "ms_per_step": ms, "elements": elements, "links": links, "containers": [n, n], "capacity": cap, "max_neighbors": maxnb,
           "dtype": "f64", "gpu_launches": launches, "steps": K, "verified": verified, "layout": stats.get("kernel"),
           "link_table_bytes": stats.get("link_table_bytes"),
           "links_per_s": links / (1e-3 * ms),
           "roofline": {"bound": "hbm", "achieved": alg / (1e-3 * ms) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (1e-3 * ms) / 1e9 / peak,
                        "traffic": (ncu_traffic("container") or {}).get("dram_bytes_per_launch") if stats.get("kernel") == 0 else None,
                        "kernel_ms": ms,
                        "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                        "note": "sweep_kernel of csrc/container.cu: 4 B per link + 28 B per element; working set %.0f MB >> L2" % (alg / 1e6)}}
    if with_e2e:
        host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in g.items()}
        for k in g:
            host[k].copy_(g[k])
        res_host = torch.empty((n, n, cap), dtype=torch.float64, pin_memory=True)
        torch.cuda.synchronize()
        del g, out
        hbox = {k: v.numpy() for k, v in host.items()}
        nbytes = sum(v.nbytes for v in hbox.values())
        t0 = time.perf_counter()
        grid.loadCells(hbox)
        grid.dev.step(capi.KERNEL_CONTAINER, K)
        grid.dev.save({"values": res_host.numpy()})
        e_s = time.perf_counter() - t0
        res["e2e"] = {"value": 1e-9 * elements * K / e_s, "unit": "Gelements/s", "h2d_bytes_per_step": nbytes / K,
                      "d2h_bytes_per_step": res_host.numpy().nbytes / K, "ms_per_run": 1e3 * e_s,
                      "what": "ContainerGrid.loadCells (pinned host arrays, %.2f GB) -> link resolution -> %d sweeps -> temperatures "
                              "back to the host" % (nbytes / 1e9, K)}
    if not args.no_cpu:
        try:   # the reference's own OpenMPSimulator over ContainerCell containers on the tile
            from oracle import oracle_py
            _, st = oracle_py.run_ref_container(box, 10, n_dims=2, torus=True, omp=True, threads=os.cpu_count(), want_output=False)
            res["cpu_baseline"] = {"value": st["geups_compute"], "unit": "Gelements/s", "cores": os.cpu_count(), "kind": "reference",
                                   "sample": "the %d^2 tile (%d elements) x 10 steps, reference OpenMPSimulator over "
                                             "ContainerCell<MeshElement, 16>, TimeCompute interval" % (tile, st["elements"])}
        except Exception as e:
            res["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (e,)}
    del grid
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200geo", choices=["b200geo", "reference"])
    ap.add_argument("--workload", default="jacobi27", choices=sorted(WORKLOADS))
    ap.add_argument("--ghost", type=int, default=0, help="ghost zone width = steps between halo exchanges (N > 1); "
                    "0 = the workload's temporal blocking depth")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary workloads (configs 0, 1, 3)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of the e2e result")
    ap.add_argument("--no-stream", action="store_true", help="skip the streamed e2e leg")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the b200geo hot path has no CPU fallback")
    torch.cuda.set_device(local)
    affinity = os.sched_getaffinity(0)
    numa = numa_bind(local) if not args.no_numa else {"bound": False}
    if world > 1:
        # the halo transfer must get SMs while the interior update is running: high-priority NCCL stream
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    t_start = time.perf_counter()

    main_res = bench_device(args.workload, args, rank, world, dist if world > 1 else None, torch)
    others = []
    if not args.no_others:
        for w in ["jacobi7", "lbm", "gol", "jacobi7_128"]:
            if w == args.workload:
                continue
            if world > 1 and w in ("gol", "jacobi7_128", "jacobi7"):
                continue
            r = bench_device(w, args, rank, world, dist if world > 1 else None, torch,
                             with_e2e=(w in ("gol", "jacobi7_128")), with_clocks=False)
            others.append({k: r[k] for k in ("workload", "value", "ms_per_step", "roofline", "dtype", "dims_per_gpu",
                                             "global_dims", "gpu_launches") if k in r} | ({"e2e": r["e2e"]} if "e2e" in r else {}))

    if not args.no_others and world in (1, 8):
        try:
            others.append(bench_nbody(args, rank, world, dist if world > 1 else None, torch))
        except Exception as e:  # secondary workload: never take the headline line down with it
            others.append({"workload": "nbody", "error": repr(e)})

    if not args.no_others and world == 1:
        try:
            others.append(bench_container(args, torch))
        except Exception as e:  # secondary workload: never take the headline line down with it
            others.append({"workload": "container", "error": repr(e)})

    os.sched_setaffinity(0, affinity)       # the CPU legs below use all host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            _, cpu = run_cpu_reference(args.workload, 10, 1, full_steps=4)
        except Exception as e:  # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "GLUPS", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}

    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            gpu_ref = gpu_reference(args.workload, args, torch)
        except Exception as e:  # a comparator, never required for the line
            gpu_ref = {"error": repr(e)}

    cpp = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpp = e2e_cpp(args)
        except Exception as e:  # a comparator, never required for the line
            cpp = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": "GLUPS (giga lattice updates/s)", "value": main_res["value"], "unit": "GLUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": main_res["dtype"],
            "data": "synthetic",
            "config": {"workload": DESCRIPTION[args.workload], "cell": main_res["model"],
                       "dims_per_gpu": main_res["dims_per_gpu"], "global_dims": main_res["global_dims"],
                       "partition": "z-slabs x%d" % world, "ghost_width": main_res["ghost_width"], "numa": numa,
                       "halo_bytes_per_exchange_per_rank": main_res["halo_bytes_per_exchange"],
                       "l2": "grids (2 x %.1f GB per GPU) >> 126 MB L2, no flush needed" %
                             (float(np.prod(main_res["dims_per_gpu"])) * {"f64": 8, "f32": 96, "u8": 1}[main_res["dtype"]] / 1e9)},
            "roofline": main_res["roofline"], "e2e": main_res.get("e2e"), "gpu_launches": main_res["gpu_launches"],
            "clocks": main_res["clocks"],
        }
        if main_res.get("verified") is not None:
            line["verified"] = main_res["verified"]
        if main_res.get("host_link") is not None:
            line["host_link"] = main_res["host_link"]
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        if cpp is not None:
            line["e2e_cpp"] = cpp
        # the secondary workloads' headline numbers as flat top-level scalars (the full records follow in "others")
        for r in others:
            w = r.get("workload")
            if w and "value" in r:
                line["%s_%s" % (w, {"nbody": "gparticles_per_s", "container": "gelements_per_s"}.get(w, "glups"))] = r["value"]
                if "roofline" in r:
                    line["%s_roofline_frac" % w] = r["roofline"].get("frac")
                if "e2e" in r:
                    line["%s_e2e" % w] = r["e2e"].get("value")
        if others:
            line["others"] = others
        line["wall_s"] = time.perf_counter() - t_start
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
