#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1p_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r1p_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/prof_r1p_tb27 python tools/tune.py jacobi27 jacobi.tb=2 > gpurun_out/r1p_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/prof_r1p_tb7 python tools/tune.py jacobi7 jacobi.tb=2 > gpurun_out/r1p_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbm_kernel -s 3 -c 1 -o gpurun_out/prof_r1p_lbm python tools/tune.py lbm > gpurun_out/r1p_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gol_kernel -s 3 -c 1 -o gpurun_out/prof_r1p_gol python tools/tune.py gol > gpurun_out/r1p_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/prof_r1p_nbody python tools/nbody_bench.py 108 2 f4 > gpurun_out/r1p_ncu5.log 2>&1
python tools/tune.py jacobi7_128 jacobi.tb=1,2,3,4 jacobi.tb_rows=32,33 > gpurun_out/r1p_tune128.log 2>&1
cat gpurun_out/r1p_tune128.log
ls -la gpurun_out/*.ncu-rep | tail -8
