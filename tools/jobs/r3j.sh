#!/bin/bash
# GPU job r3j: resident kernel without the per-thread fence; n-body: pair re-bin + 224 threads; bench N = 1 with live traffic.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_nbody_gpu.py -q -m gpu -x -k "resident or nbody" > gpurun_out/r3j_pytest.log 2>&1; tail -4 gpurun_out/r3j_pytest.log
timeout 300 python tools/tune.py jacobi7_128 jacobi.resident=0,1 jacobi.tb=1 2>&1 | tail -2 | tee gpurun_out/r3j_tune.log
for t in 256 224 192; do echo "threads $t"; NBODY_THREADS=$t timeout 300 python tools/nbody_bench.py 108 10 f4 2>&1 | tail -1; done | tee gpurun_out/r3j_nbody.log
echo "f8"; timeout 300 python tools/nbody_bench.py 64 10 f8 2>&1 | tail -1 | tee -a gpurun_out/r3j_nbody.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-others 2> gpurun_out/r3j_bench.err | grep '^{' > gpurun_out/r3j_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3j_bench.json"))
print("value %.1f frac %.3f e2e %.1f traffic %s dram_frac %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["roofline"].get("traffic"), d["roofline"].get("dram_frac")))
print(d["roofline"].get("traffic_source")); print("e2e_cpp", json.dumps(d.get("e2e_cpp"))[:400])
PY
tail -3 gpurun_out/r3j_bench.err
