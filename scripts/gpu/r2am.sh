#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 300 -k fused_aux 2>&1 | tail -15
