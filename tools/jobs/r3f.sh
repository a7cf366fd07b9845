#!/bin/bash
# GPU job r3f: the whole GPU suite with B200Stepper and the box fast path, the generic SoA bench, bench.py with the new defaults.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -rfEs > gpurun_out/r3f_pytest.log 2>&1; tail -6 gpurun_out/r3f_pytest.log
timeout 300 tests/facade/_bin/stepper_test | tail -12
timeout 300 tests/facade/_bin/generic_test 2>&1 | grep -i "stepper"
timeout 300 tests/facade/_bin/generic_soa_bench --bench | tee gpurun_out/r3f_generic_soa_bench.jsonl
timeout 300 tests/facade/_bin/generic_test --bench | tee gpurun_out/r3f_generic_bench.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r3f_bench.err | grep '^{' > gpurun_out/r3f_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3f_bench.json"))
print("value %.1f GLUPS; frac %.3f; e2e %.1f" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
print("gpu_ref", d.get("gpu_reference"))
print({k: v for k, v in d.items() if k.endswith(("_glups", "_frac", "_e2e", "_per_s"))}, "wall", d.get("wall_s"))
PY
