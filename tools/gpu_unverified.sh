#!/bin/bash
# First GPU job of the next round: run what was written after round 1's GPU budget was spent (tests/test_zz_unverified_gpu.py),
# then the throughput of the generic SoA path beside the word-sliced one.
mkdir -p gpurun_out
B200GEO_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_zz_unverified_gpu.py -q -m gpu -rfEs > gpurun_out/unverified_pytest.log 2>&1; tail -15 gpurun_out/unverified_pytest.log
timeout 300 tests/facade/_bin/generic_soa_test --bench | tee gpurun_out/generic_soa_bench.jsonl
timeout 300 tests/facade/_bin/generic_test --bench | tee gpurun_out/generic_bench.jsonl
# the streamed e2e leg of bench.py (verified against the plain schedule inside the run)
timeout 900 python bench.py --no-others 2> gpurun_out/bench_streamed.err | grep '^{' > gpurun_out/bench_streamed.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_streamed.json"))
print("value %.1f GLUPS; e2e %s" % (d["value"], json.dumps(d["e2e"])[:900]))
PY
timeout 600 python tools/stream_bench.py --chunks 8,16,32 | tee gpurun_out/stream_bench.jsonl
