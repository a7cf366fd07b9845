#!/bin/bash
# GPU job r3b: tile shapes of the temporal-blocked Jacobi kernel (4 cells per lane, 64-byte-aligned cores):
# parity of the new shapes, throughput per shape, DRAM traffic per shape (ncu, three metrics).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "temporal_blocking" > gpurun_out/r3b_pytest.log 2>&1; tail -3 gpurun_out/r3b_pytest.log
timeout 600 python tools/tune.py jacobi27 jacobi.tb=2 jacobi.tb_rows=33,40,41,42,43,44,32,64 > gpurun_out/r3b_tune27.log 2>&1; cat gpurun_out/r3b_tune27.log
timeout 600 python tools/tune.py jacobi27 jacobi.tb=3 jacobi.tb_rows=33,41,42,31 >> gpurun_out/r3b_tune27.log 2>&1; tail -4 gpurun_out/r3b_tune27.log
timeout 600 python tools/tune.py jacobi7 jacobi.tb=4 jacobi.tb_rows=33,40,41,42,31 > gpurun_out/r3b_tune7.log 2>&1; cat gpurun_out/r3b_tune7.log
timeout 600 python tools/tune.py jacobi7 jacobi.tb=3 jacobi.tb_rows=33,40,41,42,31 >> gpurun_out/r3b_tune7.log 2>&1; tail -5 gpurun_out/r3b_tune7.log
for rows in 33 40 41 42; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --clock-control none -k regex:jacobi_tb -s 2 -c 1 --csv --log-file gpurun_out/r3b_ncu_$rows.csv python tools/few_launches.py jacobi27 jacobi.tb=2 jacobi.tb_rows=$rows > /dev/null 2>&1
  echo "rows $rows"; grep -E "dram__bytes|gpu__time|lts__" gpurun_out/r3b_ncu_$rows.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
