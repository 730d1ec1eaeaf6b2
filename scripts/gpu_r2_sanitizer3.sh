# round 2: compute-sanitizer memcheck over the two-GPU path (slab decomposition, halo pack / unpack, migration, NCCL send / recv) through the host mirror:
# one process drives both GPUs (WELDFORM_GPUS=2)
tag=${1:-r2_sanitizer3}
B=$PWD/weldform_b200/host/bin
mkdir -p /tmp/san3 && cd /tmp/san3
WELDFORM_GPUS=2 WELDFORM_SYNC=tagged WELDFORM_MAX_STEPS=25 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 $B/ref_Compression_thermal > $OLDPWD/gpurun_out/sanitizer_2gpu_thermal_$tag.log 2>&1
echo "== Compression_thermal on 2 GPUs: rc=$? $(grep -c 'Solving is finished' $OLDPWD/gpurun_out/sanitizer_2gpu_thermal_$tag.log) finished; $(grep 'ERROR SUMMARY' $OLDPWD/gpurun_out/sanitizer_2gpu_thermal_$tag.log | tail -1)"
WELDFORM_GPUS=2 WELDFORM_SYNC=full WELDFORM_MAX_OUTPUTS=2 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 $B/ref_Contact_Compression_half_ghost > $OLDPWD/gpurun_out/sanitizer_2gpu_kd_contact_$tag.log 2>&1
echo "== Contact_Compression_half_ghost on 2 GPUs: rc=$? $(grep -c 'Solving is finished' $OLDPWD/gpurun_out/sanitizer_2gpu_kd_contact_$tag.log) finished; $(grep 'ERROR SUMMARY' $OLDPWD/gpurun_out/sanitizer_2gpu_kd_contact_$tag.log | tail -1)"
cd $OLDPWD
for f in gpurun_out/sanitizer_2gpu_*_$tag.log; do grep -v "^$" $f | grep -i "COMPUTE-SANITIZER\|ERROR SUMMARY\|Solving is finished\|Invalid\|BRIDGE\|weldform_b200:\|=========" | head -60 > $f.short; mv $f.short $f; done
