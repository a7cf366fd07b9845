#!/bin/bash
# The round-end sequence on one GPU: all GPU tests, smoke, both bench arms, the launch list of the bench command, one
# --set full capture of the headline kernel. Output under gpurun_out/final_*; the summaries go to profiles/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -rfEs > gpurun_out/final_pytest.log 2>&1; tail -5 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; cut -c1-300 gpurun_out/final_bench_reference.json
timeout 1200 python bench.py --steps 20 --warmup 5 2> gpurun_out/final_bench.err | grep '^{' > gpurun_out/final_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_bench.json")); e = d["e2e"]; r = d["roofline"]
print("value %.1f GLUPS; frac %.3f dram_frac %s traffic %s; e2e %.1f (%s)" % (d["value"], r["frac"], r.get("dram_frac"), r.get("traffic"), e["value"], e["schedule"][:30]))
print("verified", d.get("verified", {}).get("ok"), e.get("verified"), "| cpu", d.get("cpu_baseline", {}).get("value"), "| gpu_ref", (d.get("gpu_reference") or {}).get("value"), (d.get("gpu_reference") or {}).get("b200geo_same_dims"))
print("e2e_cpp", (d.get("e2e_cpp") or {}).get("value"), {k: (v.get("value") if isinstance(v, dict) else v) for k, v in (d.get("e2e_cpp") or {}).items()})
print({k: v for k, v in d.items() if k.endswith(("_glups", "_frac", "_e2e", "_per_s"))}, "wall", d.get("wall_s"))
PY
tail -3 gpurun_out/final_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-verify --no-stream > gpurun_out/final_bench_under_ncu.log 2>&1; wc -l gpurun_out/final_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/final_tb27_full python tools/few_launches.py jacobi27 jacobi.tb=2 > /dev/null 2>&1; ls -la gpurun_out/final_tb27_full.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/final_tb7_full python tools/few_launches.py jacobi7 jacobi.tb=4 --sweeps 16 > /dev/null 2>&1; ls -la gpurun_out/final_tb7_full.ncu-rep
