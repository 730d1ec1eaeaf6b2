# as gpu_r2_sanitizer3.sh, with CUDA API error reports off: NCCL's own set-up calls cudaDeviceEnablePeerAccess twice and reads the resulting
# cudaErrorPeerAccessAlreadyEnabled back (234 such reports, all inside libnccl.so); what is asked here is memory errors of the kernels
tag=${1:-r2_sanitizer4}
B=$PWD/weldform_b200/host/bin
mkdir -p /tmp/san4 && cd /tmp/san4
WELDFORM_GPUS=2 WELDFORM_SYNC=tagged WELDFORM_MAX_STEPS=25 timeout 300 compute-sanitizer --tool memcheck --report-api-errors no --print-limit 40 $B/ref_Compression_thermal > $OLDPWD/gpurun_out/sanitizer_2gpu_thermal_$tag.log 2>&1
echo "== Compression_thermal on 2 GPUs: rc=$? $(grep -c 'Solving is finished' $OLDPWD/gpurun_out/sanitizer_2gpu_thermal_$tag.log) finished; $(grep 'ERROR SUMMARY' $OLDPWD/gpurun_out/sanitizer_2gpu_thermal_$tag.log | tail -1)"
cd $OLDPWD
grep -v "^$" gpurun_out/sanitizer_2gpu_thermal_$tag.log | grep -i "COMPUTE-SANITIZER\|ERROR SUMMARY\|Solving is finished\|Invalid\|BRIDGE\|weldform_b200:\|=========" | head -80 > gpurun_out/s.short; mv gpurun_out/s.short gpurun_out/sanitizer_2gpu_thermal_$tag.log
