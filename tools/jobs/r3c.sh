#!/bin/bash
# GPU job r3c: where does the 27-point kernel's DRAM read over-fetch (11.0 GB vs 8.6 GB compulsory) come from?
# L2 promotion of the TMA loads, z-chunk length, CTA order; one --set full capture of the default.
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum
run() { # name, tuning...
  name=$1; shift
  timeout 300 ncu --metrics $M --clock-control none -k regex:jacobi -s 2 -c 1 --csv --log-file gpurun_out/r3c_ncu_$name.csv python tools/few_launches.py jacobi27 "$@" > /dev/null 2>&1
  echo "== $name: $@"; grep -E "dram__bytes|gpu__time|lts__" gpurun_out/r3c_ncu_$name.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ';'; echo
}
run promo0 jacobi.tb=2 jacobi.tb_promo=0
run promo2 jacobi.tb=2 jacobi.tb_promo=2
run z256 jacobi.tb=2 jacobi.tb_zchunk=256
run z512 jacobi.tb=2 jacobi.tb_zchunk=512
run z64 jacobi.tb=2 jacobi.tb_zchunk=64
run rast4 jacobi.tb=2 jacobi.tb_raster=4
run rast8 jacobi.tb=2 jacobi.tb_raster=8
run rast37 jacobi.tb=2 jacobi.tb_raster=37
run single jacobi.tb=1
timeout 600 python tools/tune.py jacobi27 jacobi.tb=2 jacobi.tb_promo=0,2,3 > gpurun_out/r3c_tune.log 2>&1
timeout 600 python tools/tune.py jacobi27 jacobi.tb=2 jacobi.tb_zchunk=64,128,256,512 >> gpurun_out/r3c_tune.log 2>&1
timeout 600 python tools/tune.py jacobi27 jacobi.tb=2 jacobi.tb_raster=0,2,4,8,16,37 >> gpurun_out/r3c_tune.log 2>&1
timeout 600 python tools/tune.py jacobi7 jacobi.tb=4 jacobi.tb_raster=0,4,8 jacobi.tb_zchunk=128,256 >> gpurun_out/r3c_tune.log 2>&1
cat gpurun_out/r3c_tune.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/r3c_tb27_full python tools/few_launches.py jacobi27 jacobi.tb=2 > /dev/null 2>&1; ls -la gpurun_out/r3c_tb27_full.ncu-rep
