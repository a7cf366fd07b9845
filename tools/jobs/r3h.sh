#!/bin/bash
# GPU job r3h: GPU suite (stepper, checkpoints, read-ahead), bench.py incl. the C++ e2e leg.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -rfEs > gpurun_out/r3h_pytest.log 2>&1; tail -6 gpurun_out/r3h_pytest.log
for mode in box rows; do timeout 300 tests/facade/_bin/e2e_bench 1024 20 1 $mode; done | tee gpurun_out/r3h_e2e_cpp.jsonl | cut -c1-330
timeout 300 tests/facade/_bin/e2e_bench 512 20 1 box --bov /tmp/b200bov | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r3h_bench.err | grep '^{' > gpurun_out/r3h_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3h_bench.json"))
e = d["e2e"]
print("value %.1f GLUPS; frac %.3f; e2e %.1f (%s)" % (d["value"], d["roofline"]["frac"], e["value"], e["schedule"][:30]))
print("e2e_cpp", json.dumps(d.get("e2e_cpp"))[:900])
print({k: v for k, v in d.items() if k.endswith(("_glups", "_frac", "_e2e", "_per_s"))}, "wall", d.get("wall_s"))
PY
tail -3 gpurun_out/r3h_bench.err
