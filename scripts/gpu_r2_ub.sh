tag=${1:-r2_ub}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/ubench/cell_sync.cu -o /tmp/cell_sync && /tmp/cell_sync > gpurun_out/cell_sync_$tag.txt 2>&1; cat gpurun_out/cell_sync_$tag.txt
python -m pytest tests/test_host_api.py -m gpu -q -k "pins_temperature" 2>&1 | tail -3
