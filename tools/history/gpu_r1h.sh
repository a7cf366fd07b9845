#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nbody_gpu.py -x -q -m gpu > gpurun_out/r1h_pytest.log 2>&1
tail -15 gpurun_out/r1h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
