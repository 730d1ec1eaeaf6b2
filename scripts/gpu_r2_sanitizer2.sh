# round 2: compute-sanitizer memcheck over the host mirror's bridge paths (registered subsets, fill kernels, full up / downloads) with the reference's own drivers
tag=${1:-r2_sanitizer2}
B=$PWD/weldform_b200/host/bin
mkdir -p /tmp/san2 && cd /tmp/san2
for mode in push tagged full; do
  WELDFORM_SYNC=$mode WELDFORM_MAX_STEPS=25 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 $B/ref_Compression_thermal > $OLDPWD/gpurun_out/sanitizer_mirror_${mode}_$tag.log 2>&1
  echo "== Compression_thermal $mode: $(grep -c 'Solving is finished' $OLDPWD/gpurun_out/sanitizer_mirror_${mode}_$tag.log) finished; $(grep 'ERROR SUMMARY' $OLDPWD/gpurun_out/sanitizer_mirror_${mode}_$tag.log | tail -1)"
done
WELDFORM_SYNC=full WELDFORM_MAX_OUTPUTS=2 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 $B/ref_Contact_Compression_half_ghost > $OLDPWD/gpurun_out/sanitizer_mirror_kd_contact_$tag.log 2>&1
echo "== Contact_Compression_half_ghost: $(grep -c 'Solving is finished' $OLDPWD/gpurun_out/sanitizer_mirror_kd_contact_$tag.log) finished; $(grep 'ERROR SUMMARY' $OLDPWD/gpurun_out/sanitizer_mirror_kd_contact_$tag.log | tail -1)"
cd $OLDPWD
for f in gpurun_out/sanitizer_mirror_*_$tag.log; do grep -v "^$" $f | grep -i "COMPUTE-SANITIZER\|ERROR SUMMARY\|Solving is finished\|Invalid\|BRIDGE\|weldform_b200:" > $f.short; mv $f.short $f; done
