#!/bin/bash
# GPU job r3l (2 GPUs): e2e legs at N = 2 (streamed run with ghost zones as wide as the run), multi-GPU tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py -q -m gpu -x -rs > gpurun_out/r3l_pytest.log 2>&1; tail -4 gpurun_out/r3l_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-others 2> gpurun_out/r3l_n2.err | grep '^{' > gpurun_out/r3l_n2.json
tail -3 gpurun_out/r3l_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3l_n2.json")); e = d["e2e"]
print("N=2 value %.1f e2e %.1f (%s) ms %.1f" % (d["value"], e["value"], e["schedule"][:60], e["ms_per_run"]))
print("   plain:", e.get("plain_schedule"), "| streamed:", json.dumps(e.get("streamed_schedule"))[:700])
print("   verified:", d.get("verified", {}).get("per_rank"), e.get("verified"), e.get("how", "")[:400])
PY
timeout 300 tests/facade/_bin/e2e_bench 1024 20 2 box | cut -c1-250
