#!/bin/bash
# the round-end sequence on one GPU: all GPU tests, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full_pytest.log 2>&1
tail -4 gpurun_out/full_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
