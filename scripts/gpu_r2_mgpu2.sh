tag=${1:-r2_mgpu2}
python -m pytest tests/test_multi_gpu.py -m gpu -q -rA 2>&1 | tail -40 > gpurun_out/pytest_mgpu_$tag.log
grep -E "PASSED|FAILED|ERROR|passed|failed" gpurun_out/pytest_mgpu_$tag.log | tail -30
