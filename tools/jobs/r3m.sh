#!/bin/bash
# GPU job r3m (8 GPUs): weak scaling and e2e at N = 8 with the driver's flags, multi-GPU parity on 4 and 8 GPUs, C++ slab group e2e.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; free -g | head -2; nproc
timeout 600 python -m pytest tests/test_multigpu.py -q -m gpu -x -rs > gpurun_out/r3m_pytest.log 2>&1; tail -3 gpurun_out/r3m_pytest.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-others 2> gpurun_out/r3m_n$n.err | grep '^{' > gpurun_out/r3m_n$n.json
tail -2 gpurun_out/r3m_n$n.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu 2> gpurun_out/r3m_n1.err | grep '^{' > gpurun_out/r3m_n1.json
python - <<'PY'
import json
base = None
for n in (1, 4, 8):
    try:
        d = json.load(open("gpurun_out/r3m_n%d.json" % n))
    except Exception as e:
        print(n, "no line", e); continue
    e = d["e2e"]
    if n == 1: base = d
    print("N=%d value %.1f (eff %.3f) e2e %.1f (%s) ms %.1f | other schedule: %s" % (n, d["value"], d["value"] / (n * base["value"]) if base else 0, e["value"], e["schedule"][:28], e["ms_per_run"], json.dumps(e.get("plain_schedule") or e.get("streamed_schedule"))[:260]))
    print("   verified:", d.get("verified", {}).get("per_rank"), e.get("verified"))
PY
for s in 1 8; do timeout 300 tests/facade/_bin/e2e_bench 1024 20 $s box | cut -c1-200; done
