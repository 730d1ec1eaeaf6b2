# round 2, third GPU call: re-run of what failed / changed, slot-sorted lists under the contracted kernels (A/B), e2e through the C++ API again
tag=${1:-r2_v2}
mkdir -p gpurun_out/drivers_$tag
export WF_PARITY_REPORT=$PWD/gpurun_out/parity_report_$tag.jsonl
rm -f $WF_PARITY_REPORT
python -m pytest tests -m gpu -q --deselect tests/test_drivers_gpu.py 2>&1 | tail -40 > gpurun_out/pytest_$tag.log
tail -8 gpurun_out/pytest_$tag.log
WF_KEEP_DRIVER_OUT=$PWD/gpurun_out/drivers_$tag timeout 1500 python -m pytest tests/test_drivers_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_drivers_$tag.log
tail -6 gpurun_out/pytest_drivers_$tag.log
find gpurun_out/drivers_$tag -name "test06_*.csv" -delete
for t in 1024 5120 1040; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tuning $t > gpurun_out/bench_${tag}_t$t.json 2> gpurun_out/bench_${tag}_t$t.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}_t$t.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("tuning $t: %.2f ms/step, accel %.2f rate %.2f post %.2f nb_list %.2f ms, e2e %.3g (%s), checksum %s" % (d["ms_per_step"], k["accel"]["ms_per_launch"], k["rate_stress"]["ms_per_launch"], k["post"]["ms_per_launch"], k["nb_list"]["ms_per_launch"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}, d["config"]["state_checksum"]["x"]))
PY
done
ncu --set full --clock-control none --import-source on -k regex:"k_accel_fm|k_post_fm|k_rate_fm|k_nb_list" -s 6 -c 8 -o gpurun_out/prof_${tag}_fm python bench.py --steps 3 --warmup 3 --max-spinup 20 --no-cpu-baseline --python-e2e --tuning 1024 > gpurun_out/bench_ncu_$tag.log 2>&1
ls -la gpurun_out | tail -8
