#!/bin/bash
# GPU job r3e (2 GPUs): the NCCL halo path's parity on real devices, weak scaling 1 -> 2 with the driver's flags
# (--warmup 5 --steps 20: the misaligned-round case), per-rank oracle verification; n-body v2 without the 27-fold unroll.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -x -rs > gpurun_out/r3e_pytest.log 2>&1; tail -4 gpurun_out/r3e_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-others --no-cpu 2> gpurun_out/r3e_n1.err | grep '^{' > gpurun_out/r3e_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-others 2> gpurun_out/r3e_n2.err | grep '^{' > gpurun_out/r3e_n2.json
tail -3 gpurun_out/r3e_n2.err
python - <<'PY'
import json
a = json.load(open("gpurun_out/r3e_n1.json")); b = json.load(open("gpurun_out/r3e_n2.json"))
print("N=1 %.1f GLUPS e2e %.1f | N=2 %.1f GLUPS e2e %.1f | efficiency %.3f | launches %s / %s" % (
    a["value"], a["e2e"]["value"], b["value"], b["e2e"]["value"], b["value"] / (2 * a["value"]), a["gpu_launches"], b["gpu_launches"]))
print("verified", a.get("verified"), b.get("verified"))
PY
NBODY_RUN=16 timeout 300 python tools/nbody_bench.py 108 10 f4 2>&1 | tail -1 | tee gpurun_out/r3e_nbody.log
