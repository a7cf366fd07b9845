#!/bin/bash
# GPU job r5f: the whole GPU suite and smoke() on the final tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r5f_pytest.log 2>&1; tail -3 gpurun_out/r5f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
