#!/bin/bash
# GPU job r6b: ContainerCell path with the sliced-ELLPACK link table: the whole GPU suite, ncu --set full of the new
# sweep / window-sort / resolve kernels, then the default bench.py run (with the container leg among the others)
mkdir -p gpurun_out
timeout 260 python -m pytest tests -m gpu -x -q > gpurun_out/r6b_pytest.log 2>&1; tail -4 gpurun_out/r6b_pytest.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"window_sort_kernel|resolve_kernel|sweep_kernel" -c 4 -o gpurun_out/r6b_container_full python tools/container_bench.py --steps 20 --no-cpu --no-e2e --no-verify > /dev/null 2>&1; ls -la gpurun_out/r6b_container_full.ncu-rep
timeout 280 python bench.py > gpurun_out/r6b_bench.json 2> gpurun_out/r6b_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r6b_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in d if k not in ("others", "config", "clocks", "e2e_cpp", "gpu_reference", "host_link")})
    print([o for o in d.get("others", []) if o.get("workload") == "container"])
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -3 gpurun_out/r6b_bench.err
