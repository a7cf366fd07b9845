#!/bin/bash
# 8-GPU box, final build of round 1: multi-GPU parity (NCCL ranks at world 4, slab groups and the C++ striping façade
# over 8 devices), weak scaling of bench.py at N = 8 and 4, single-process group at N = 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py -x -q -m gpu -k "4 or group" > gpurun_out/s8_pytest.log 2>&1; tail -3 gpurun_out/s8_pytest.log
tests/facade/_bin/striping_test 2>&1 | tail -3
for wl in jacobi27 lbm; do
  timeout 300 python tools/group_bench.py $wl 8 2>&1 | tail -1 | tee -a gpurun_out/s8_group_bench.jsonl
done
for n in 8 4; do
  for wl in jacobi27 lbm; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 100 --warmup 6 --workload $wl --no-cpu $( [ $wl = jacobi27 ] && [ $n = 8 ] && echo "" || echo --no-others ) 2> gpurun_out/s8_n${n}_$wl.err | grep '^{' > gpurun_out/s8_n${n}_$wl.json
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8_n${n}_$wl.json"))
    print("N=$n $wl: value %.1f %s  ms/step %.4f  kernel_ms/launch %.4f  e2e %.1f  launches %d" % (d["value"], d["unit"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
    for o in d.get("others", []):
        print("   ", o.get("workload"), {k: o.get(k) for k in ("value","ms_per_step","error")})
except Exception as e:
    print("N=$n $wl failed", e); print(open("gpurun_out/s8_n${n}_$wl.err").read()[-1500:])
PY
  done
done
