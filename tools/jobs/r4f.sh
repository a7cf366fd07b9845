#!/bin/bash
# GPU job r4f: fused LBM kernel: DRAM traffic and time per launch for L2 hints, z chunks, tile rows (ncu byte counters, one launch each)
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
for v in "lbm.tb_hints=0" "lbm.tb_hints=1" "lbm.tb_hints=2" "lbm.tb_hints=3" "lbm.tb_hints=3 lbm.tb_zchunk=512" "lbm.tb_hints=3 lbm.tb_zchunk=32" "lbm.tb_hints=3 lbm.tb_rows=16" "lbm.tb_hints=3 lbm.tb_rows=8" "lbm.tb_hints=3 lbm.tb_promo=3"; do
  echo "== $v"
  timeout 300 ncu --metrics $M --clock-control none -k regex:lbm_tb2 -s 1 -c 1 --csv python tools/few_launches.py lbm lbm.tb=2 $v 2>/dev/null | grep lbm_tb2 | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"'
done 2>&1 | tee gpurun_out/r4f_traffic.log
timeout 900 python tools/tune.py lbm lbm.tb=2 lbm.tb_hints=0,1,2,3 lbm.tb_zchunk=64,512 2>&1 | tee gpurun_out/r4f_tune.log
