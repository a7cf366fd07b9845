#!/bin/bash
# GPU job r4d: which TMA window shapes are legal (tools/probe/tma_probe.cu), one process per case
for i in 0 1 2 3 4 5 6 7; do tools/probe/tma_probe $i; done
