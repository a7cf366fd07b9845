#!/bin/bash
# GPU job r3d: the mask-based n-body kernel (parity, throughput per run length, f32 and f64, one full ncu capture);
# Jacobi: TMA loads without L2 promotion x z-chunk length.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nbody_gpu.py -q -m gpu -x > gpurun_out/r3d_pytest.log 2>&1; tail -3 gpurun_out/r3d_pytest.log
for k in 3 0; do for r in 8 12 16; do
  [ $k = 3 ] && [ $r != 16 ] && continue
  echo "kernel $k run $r"; NBODY_KERNEL=$k NBODY_RUN=$r timeout 300 python tools/nbody_bench.py 108 10 f4 2>&1 | tail -1
done; done | tee gpurun_out/r3d_nbody.log
for r in 8 16; do echo "f8 run $r"; NBODY_RUN=$r timeout 300 python tools/nbody_bench.py 64 10 f8 2>&1 | tail -1; done | tee -a gpurun_out/r3d_nbody.log
echo "f8 old"; NBODY_KERNEL=3 timeout 300 python tools/nbody_bench.py 64 10 f8 2>&1 | tail -1 | tee -a gpurun_out/r3d_nbody.log
timeout 600 python tools/tune.py jacobi27 jacobi.tb=2 jacobi.tb_promo=0 jacobi.tb_zchunk=32,48,64,96,128 > gpurun_out/r3d_tune.log 2>&1
timeout 600 python tools/tune.py jacobi7 jacobi.tb=4 jacobi.tb_promo=0,3 jacobi.tb_zchunk=64,128,256 >> gpurun_out/r3d_tune.log 2>&1
timeout 600 python tools/tune.py jacobi27 jacobi.tb=1 >> gpurun_out/r3d_tune.log 2>&1
cat gpurun_out/r3d_tune.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 1 -c 1 -o gpurun_out/r3d_nbody_full python tools/nbody_bench.py 108 2 f4 > /dev/null 2>&1; ls -la gpurun_out/r3d_nbody_full.ncu-rep
