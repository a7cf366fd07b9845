#!/bin/bash
# Round 1 (u): 2-GPU box — slab groups over real NVLink P2P, C++ striping façade, group weak scaling
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_group_gpu.py -x -q -m gpu > gpurun_out/r1u_pytest.log 2>&1; tail -4 gpurun_out/r1u_pytest.log
tests/facade/_bin/striping_test > gpurun_out/r1u_striping.log 2>&1; echo "striping_test exit $?"; tail -20 gpurun_out/r1u_striping.log
for wl in jacobi27 lbm; do
  for n in 1 2; do
    timeout 300 python tools/group_bench.py $wl $n 2>&1 | tail -1 | tee -a gpurun_out/r1u_group_bench.jsonl
  done
done
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu -k 2 > gpurun_out/r1u_pytest_nccl.log 2>&1; tail -3 gpurun_out/r1u_pytest_nccl.log
