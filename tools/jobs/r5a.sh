#!/bin/bash
# GPU job r5a: the whole GPU suite three times in a row (flakiness check of the final build)
mkdir -p gpurun_out
for i in 1 2 3; do timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r5a_pytest_$i.log 2>&1; tail -1 gpurun_out/r5a_pytest_$i.log; done
