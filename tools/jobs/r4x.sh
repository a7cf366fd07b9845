#!/bin/bash
# GPU job r4x (4 GPUs): the headline bench at N = 4 and N = 2 with the final build (driver flags), multi-GPU parity on 4 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -q -m gpu -x -rs 2>&1 | tail -3
for n in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --no-others 2> gpurun_out/r4x_n$n.err | grep '^{' > gpurun_out/r4x_n$n.json
tail -1 gpurun_out/r4x_n$n.err | cut -c1-200
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu 2> gpurun_out/r4x_n1.err | grep '^{' > gpurun_out/r4x_n1.json
python - <<'PY'
import json
base = None
for n in (1, 2, 4):
    try:
        d = json.load(open("gpurun_out/r4x_n%d.json" % n))
    except Exception as e:
        print(n, "no line", e); continue
    e = d["e2e"]
    if n == 1: base = d
    print("N=%d value %.1f (eff %.3f) ms/step %.4f launches %s e2e %.1f (%s) verified %s %s" % (n, d["value"], d["value"] / (n * base["value"]) if base else 0, d["ms_per_step"], d.get("gpu_launches"), e["value"], e["schedule"][:28], d.get("verified", {}).get("per_rank"), e.get("verified")))
PY
