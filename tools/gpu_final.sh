#!/bin/bash
# The round-end sequence on one GPU with the final build: all GPU tests, smoke, both bench arms, launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference 2> gpurun_out/final_ref.err | grep '^{' > gpurun_out/final_bench_reference.json; cut -c1-200 gpurun_out/final_bench_reference.json
timeout 900 python bench.py 2> gpurun_out/final_bench.err | grep '^{' > gpurun_out/final_bench_n1.json; tail -2 gpurun_out/final_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_bench_n1.json"))
print("headline %s: %.1f %s, ms/step %.4f, roofline frac %.3f (dram_frac %s), e2e %.1f, launches %d, clocks %s, wall %.1f s" % (
    d["config"]["cell"], d["value"], d["unit"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), d["e2e"]["value"], d["gpu_launches"], d["clocks"], d["wall_s"]))
print("cpu_baseline", d.get("cpu_baseline"))
for o in d.get("others", []):
    print("  ", o.get("workload"), "%.2f" % o.get("value", -1), o.get("unit", "GLUPS"), "ms/step %.5f" % o.get("ms_per_step", -1), "frac %.3f" % o.get("roofline", {}).get("frac", -1), "e2e", (o.get("e2e") or {}).get("value"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/final_bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/final_bench_under_ncu.log | head -3; wc -l gpurun_out/final_launches.csv
