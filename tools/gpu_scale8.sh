#!/bin/bash
# 8-GPU box: multi-GPU parity at world 4, then the weak-scaling bench at N = 8, 4, 2 (N = 1 is measured on 1-GPU calls)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "4" > gpurun_out/scale_pytest4.log 2>&1
tail -3 gpurun_out/scale_pytest4.log
for n in 8 4 2; do
  for wl in jacobi27 lbm; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 6 --workload $wl --no-cpu $( [ $wl = lbm ] && echo --no-others || ( [ $n = 8 ] && echo "" || echo --no-others ) ) 2> gpurun_out/scale_n${n}_$wl.err | grep '^{' > gpurun_out/scale_n${n}_$wl.json
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_n${n}_$wl.json"))
    print("N=$n $wl: value %.1f %s  ms/step %.4f  kernel_ms/launch %.4f  e2e %.1f  launches %d" % (d["value"], d["unit"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
    for o in d.get("others", []):
        if o.get("workload") == "nbody": print("   nbody:", {k: o.get(k) for k in ("value","ms_per_step","error")})
except Exception as e:
    print("N=$n $wl failed", e); print(open("gpurun_out/scale_n${n}_$wl.err").read()[-1500:])
PY
  done
done
