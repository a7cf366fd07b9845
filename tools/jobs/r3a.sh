#!/bin/bash
# GPU job r3a (round 2, first): the formerly gated tests inside the whole GPU suite, bench.py with the new checker legs,
# the generic SoA path's throughput, the streamed run per chunk count.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; free -g | head -2; nproc
timeout 1200 python -m pytest tests -q -m gpu -x -rfEs > gpurun_out/r3a_pytest.log 2>&1; tail -8 gpurun_out/r3a_pytest.log
timeout 300 tests/facade/_bin/generic_soa_test --bench > gpurun_out/r3a_generic_soa_bench.jsonl 2>&1; tail -5 gpurun_out/r3a_generic_soa_bench.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r3a_bench.err | grep '^{' > gpurun_out/r3a_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3a_bench.json"))
print("value %.1f GLUPS; frac %.3f; e2e %s" % (d["value"], d["roofline"]["frac"], json.dumps(d["e2e"])[:1200]))
print("verified", d.get("verified")); print("cpu", d.get("cpu_baseline")); print("gpu_ref", d.get("gpu_reference"))
print({k: v for k, v in d.items() if k.endswith(("_glups", "_frac", "_e2e", "_per_s"))}, "wall", d.get("wall_s"))
PY
tail -5 gpurun_out/r3a_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r3a_bench_ref.json 2> gpurun_out/r3a_bench_ref.err; cut -c1-600 gpurun_out/r3a_bench_ref.json
timeout 600 python tools/stream_bench.py --chunks 8,16,32 > gpurun_out/r3a_stream_bench.jsonl 2>&1; tail -4 gpurun_out/r3a_stream_bench.jsonl
