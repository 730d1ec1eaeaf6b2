# round 2: the reference's unmodified drivers on two GPUs (WELDFORM_GPUS=2) against the reference's own tables
tag=${1:-r2_2gpu_drv}
mkdir -p gpurun_out/drv_$tag
nvidia-smi -L | head -3
WF_KEEP_DRIVER_OUT=$PWD/gpurun_out/drv_$tag timeout 1200 python -m pytest tests/test_drivers_gpu.py -m gpu -q -k two_gpus 2>&1 | grep -v "^  \|^$" | cut -c1-400 | tail -80 > gpurun_out/pytest_$tag.log
tail -40 gpurun_out/pytest_$tag.log
find gpurun_out/drv_$tag -name "*.csv" -size +100k -delete
