#!/bin/bash
# GPU job r4n (2 GPUs): B200PatchLink between steppers on different GPUs (stepper_test), the slab simulators on 2 GPUs, multi-GPU parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 tests/facade/_bin/stepper_test 2>&1 | tee gpurun_out/r4n_stepper_test.log | tail -9
timeout 600 tests/facade/_bin/striping_test 2>&1 | tail -3
timeout 900 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py -q -m gpu -x -rs 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload lbm --no-others 2> gpurun_out/r4n_lbm_n2.err | grep '^{' > gpurun_out/r4n_lbm_n2.json; tail -2 gpurun_out/r4n_lbm_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r4n_lbm_n2.json").read().strip().splitlines()[-1])
print("LBM N=2 value", d["value"], "ghost", d["config"].get("ghost_width"), "launches", d.get("gpu_launches"), "verified", d.get("verified", {}).get("per_rank"), "e2e", d["e2e"]["value"], d["e2e"].get("schedule", "")[:40])
PY
