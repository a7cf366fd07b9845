#!/bin/bash
# Round 1 (2a): full suite with dependent launches (Jacobi small grids, packed GoL), generic-path throughput,
# memcheck of one small case per kernel family
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
timeout 300 tests/facade/_bin/generic_test --bench 2>&1 | tee gpurun_out/r2a_generic_bench.jsonl
timeout 600 python bench.py --workload gol --no-others --no-cpu 2>/dev/null | grep '^{' > gpurun_out/r2a_bench_gol.json
timeout 600 python bench.py --workload jacobi7_128 --no-others --no-cpu 2>/dev/null | grep '^{' > gpurun_out/r2a_bench_j128.json
python - <<'PY'
import json
for n in ("gol", "j128"):
    d = json.load(open("gpurun_out/r2a_bench_%s.json" % n))
    print("%s: %.1f %s, ms/step %.5f, e2e %.1f" % (n, d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"]))
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "jacobi_bit_exact or lbm_bit_exact or gol_bit_exact or z_chunks" > gpurun_out/r2a_memcheck.log 2>&1; echo "memcheck exit $?"; tail -6 gpurun_out/r2a_memcheck.log
