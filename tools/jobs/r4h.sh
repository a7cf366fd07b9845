#!/bin/bash
# GPU job r4h: the whole GPU suite with the fused LBM path on by default (one GPU, slab groups, streamed run), then the LBM bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r4h_pytest.log 2>&1; tail -5 gpurun_out/r4h_pytest.log
timeout 600 python bench.py --workload lbm --steps 20 --warmup 5 --no-others 2> gpurun_out/r4h_lbm.err | grep '^{' > gpurun_out/r4h_lbm.json; tail -3 gpurun_out/r4h_lbm.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r4h_lbm.json").read().strip().splitlines()[-1])
print("value", d["value"], "roofline", {k: d["roofline"].get(k) for k in ("frac", "dram_frac", "traffic", "kernel_ms", "sweeps_per_launch")})
print("e2e", d.get("e2e"), "verified", d.get("verified"))
PY
