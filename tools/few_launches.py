#!/usr/bin/env python
"""A handful of launches of one kernel variant, as the target of an ncu capture:
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:<kernel> -s 2 -c 1 \
       python tools/few_launches.py jacobi27 jacobi.tb=2 jacobi.tb_rows=40
usage: tools/few_launches.py <workload of tools/tune.py> [key=value ...] [--sweeps N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libgeodecomp_b200 import capi, synth
from libgeodecomp_b200.simulator import B200Grid
from tune import SPEC


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    sweeps = 8
    if "--sweeps" in sys.argv:
        sweeps = int(sys.argv[sys.argv.index("--sweeps") + 1])
        args = [a for a in args if a != str(sweeps)]
    model, dims, _ = SPEC[args[0]]
    for kv in args[1:]:
        k, v = kv.split("=")
        capi.set_tuning(k, int(v))
    grid = B200Grid(model, dims)
    if args[0] == "gol":
        grid.loadMember("alive", torch.randint(0, 2, dims[::-1], dtype=torch.uint8, device="cuda"), location=capi.CUDA_DEVICE)
    elif args[0] == "lbm":
        grid.loadMember("C", np.ones(dims[::-1], dtype=np.float32))
        grid.loadMember("density", np.ones(dims[::-1], dtype=np.float32))
        grid.loadMember("state", synth.lbm_states(dims[0], dims[1], dims[2], 0, dims[2]))
    else:
        grid.loadMember("temp", torch.rand(dims[::-1], dtype=torch.float64, device="cuda"), location=capi.CUDA_DEVICE)
    torch.cuda.synchronize()
    grid.dev.step(model.kernel, sweeps)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
