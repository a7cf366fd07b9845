#!/bin/bash
# GPU job r4z: whole-grid repeated-launch tests at full size (Jacobi 1024^3 fused vs single sweeps, LBM 512^3)
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -x -k "repeated" --durations=5 2>&1 | tail -12
