#!/bin/bash
# GPU job r5d (2 GPUs): LBM halos member by member and overlapped at ghost width 2 (was: packed, exchange then step): parity, N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py -q -m gpu -x -rs 2>&1 | tail -3
timeout 600 tests/facade/_bin/striping_test 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --workload lbm --no-others --no-cpu 2> gpurun_out/r5d_lbm_n2.err | grep '^{' > gpurun_out/r5d_lbm_n2.json; tail -1 gpurun_out/r5d_lbm_n2.err | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --workload lbm --no-others --no-cpu 2> gpurun_out/r5d_lbm_n1.err | grep '^{' > gpurun_out/r5d_lbm_n1.json
python - <<'PY'
import json
for n in (1, 2):
    d = json.load(open("gpurun_out/r5d_lbm_n%d.json" % n))
    e = d["e2e"]
    print("LBM N=%d value %.1f GLUPS (%.1f per GPU) ms/step %.4f ghost %s launches %s halo bytes %s e2e %.1f verified %s %s" % (n, d["value"], d["value"] / n, d["ms_per_step"], d["config"].get("ghost_width"), d.get("gpu_launches"), d["config"].get("halo_bytes_per_exchange_per_rank"), e["value"], d.get("verified", {}).get("per_rank"), e.get("verified")))
PY
