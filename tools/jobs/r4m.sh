#!/bin/bash
# GPU job r4m: fused LBM kernel: DRAM bytes per launch when all CTAs start together (one wave) vs several waves
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for shape in "512 108 32" "512 108 64" "512 108 256" "512 216 256" "512 512 64" "512 512 512"; do
  echo "== $shape (zchunk 32)"
  timeout 300 ncu --metrics $M --clock-control none -k regex:lbm_tb2w -s 1 -c 1 --csv python tools/lbm_wave_probe.py $shape 2>/dev/null | grep "lbm_tb2w\|compulsory" | awk -F'","' '{ if (NF > 3) print $(NF-2), $(NF); else print $0 }' | tr -d '"'
done 2>&1 | tee gpurun_out/r4m_waves.log
