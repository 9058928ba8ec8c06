timeout 200 python -m pytest ${1:-tests/test_gpu_pipeline.py} -x -q 2>&1 | tail -15
