#!/bin/bash
# GPU job r5e (8 GPUs; as r5c, with the overlapped halos): LBM D3Q19 512^3 per GPU, weak scaling at N = 8 and 4 (BASELINE.json configs[3]), every rank verified
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 --workload lbm --no-others --no-cpu 2> gpurun_out/r5e_lbm_n$n.err | grep '^{' > gpurun_out/r5e_lbm_n$n.json
tail -1 gpurun_out/r5e_lbm_n$n.err | cut -c1-200
done
python - <<'PY'
import json
for n in (4, 8):
    try:
        d = json.load(open("gpurun_out/r5e_lbm_n%d.json" % n))
    except Exception as e:
        print(n, "no line", e); continue
    e = d["e2e"]
    print("LBM N=%d value %.1f GLUPS (%.1f per GPU) ms/step %.4f ghost %s launches %s e2e %.1f verified %s %s" % (n, d["value"], d["value"] / n, d["ms_per_step"], d["config"].get("ghost_width"), d.get("gpu_launches"), e["value"], d.get("verified", {}).get("per_rank"), e.get("verified")))
PY
