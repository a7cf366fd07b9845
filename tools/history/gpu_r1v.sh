#!/bin/bash
# Round 1 (v): full suite, both bench arms at N = 1, launch list of the bench, ncu --set full of the new
# LBM kernel and of the 4-sweep 7-point kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r1v_pytest.log 2>&1; tail -3 gpurun_out/r1v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 2> gpurun_out/r1v_ref.err | grep '^{' > gpurun_out/r1v_bench_reference.json; cut -c1-400 gpurun_out/r1v_bench_reference.json
timeout 900 python bench.py 2> gpurun_out/r1v_bench.err | grep '^{' > gpurun_out/r1v_bench_n1.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r1v_bench_n1.json"))
print("headline %s: %.1f %s, ms/step %.4f, roofline frac %.3f (dram_frac %s), e2e %.1f, launches %d, clocks %s" % (
    d["config"]["model"], d["value"], d["unit"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
print("cpu_baseline", d.get("cpu_baseline"))
for o in d.get("others", []):
    print("  ", o.get("workload"), "%.2f" % o.get("value", -1), o.get("unit", "GLUPS"), "ms/step %.4f" % o.get("ms_per_step", -1), "frac %.3f" % o.get("roofline", {}).get("frac", -1), "e2e", (o.get("e2e") or {}).get("value"))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1v_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r1v_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbm_kernel -s 3 -c 1 -o gpurun_out/prof_r1v_lbm python tools/tune.py lbm > gpurun_out/r1v_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/prof_r1v_tb7 python tools/tune.py jacobi7 jacobi.tb=4 > gpurun_out/r1v_ncu2.log 2>&1
ls -la gpurun_out/*r1v* | tail
