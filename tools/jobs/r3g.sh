#!/bin/bash
# GPU job r3g (2 GPUs): the reworked e2e legs (in-process streamed run, NUMA binding, ghost zones as wide as the run on N > 1).
mkdir -p gpurun_out
lscpu | grep -E "NUMA|Model name|^CPU\(s\)" | head; nvidia-smi topo -m 2>/dev/null | head -8
timeout 600 python bench.py --steps 20 --warmup 5 --no-others --no-cpu 2> gpurun_out/r3g_n1.err | grep '^{' > gpurun_out/r3g_n1.json; tail -3 gpurun_out/r3g_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-others --no-cpu --no-numa 2> gpurun_out/r3g_n1_nonuma.err | grep '^{' > gpurun_out/r3g_n1_nonuma.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-others 2> gpurun_out/r3g_n2.err | grep '^{' > gpurun_out/r3g_n2.json
tail -3 gpurun_out/r3g_n2.err
python - <<'PY'
import json
for f in ("r3g_n1", "r3g_n1_nonuma", "r3g_n2"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
    except Exception as e:
        print(f, "no line:", e); continue
    e = d["e2e"]
    print(f, "value %.1f e2e %.1f (%s) ms %.1f numa %s" % (d["value"], e["value"], e["schedule"][:40], e["ms_per_run"], d["config"].get("numa")))
    print("   plain:", e.get("plain_schedule"), "| streamed:", json.dumps(e.get("streamed_schedule"))[:600])
    print("   verified:", d.get("verified", {}).get("ok"), e.get("verified"), e.get("how", "")[:300])
PY
