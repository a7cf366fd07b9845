#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "2" > gpurun_out/r1g_pytest.log 2>&1
tail -5 gpurun_out/r1g_pytest.log
for wl in jacobi27 lbm; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 6 --workload $wl --no-others --no-cpu > gpurun_out/r1g_bench_n2_$wl.json 2> gpurun_out/r1g_bench_n2_$wl.err
tail -c 1500 gpurun_out/r1g_bench_n2_$wl.json; tail -3 gpurun_out/r1g_bench_n2_$wl.err
done
timeout 600 python bench.py --steps 100 --warmup 6 --no-cpu > gpurun_out/r1g_bench_n1.json 2> gpurun_out/r1g_bench_n1.err
tail -c 3000 gpurun_out/r1g_bench_n1.json; tail -3 gpurun_out/r1g_bench_n1.err
