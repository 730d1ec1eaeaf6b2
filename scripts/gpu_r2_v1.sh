# round 2, second GPU call: element-wise parity report of the whole suite, driver runs (outputs kept for offline comparison), contracted-arithmetic A/B
tag=${1:-r2_v1}
mkdir -p gpurun_out/drivers_$tag
export WF_PARITY_REPORT=$PWD/gpurun_out/parity_report_$tag.jsonl
rm -f $WF_PARITY_REPORT
python -m pytest tests -m gpu -q --deselect tests/test_drivers_gpu.py 2>&1 | tail -60 > gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
WF_KEEP_DRIVER_OUT=$PWD/gpurun_out/drivers_$tag timeout 1500 python -m pytest tests/test_drivers_gpu.py -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_drivers_$tag.log
tail -15 gpurun_out/pytest_drivers_$tag.log
# keep only the tables the goldens pin (or may pin): 1 2 3 5 8 10 30 60 106
for f in $(find gpurun_out/drivers_$tag -name "test06_*.csv"); do
  n=$(basename $f .csv | sed 's/test06_0*//'); case " 1 2 3 5 8 10 30 60 106 " in *" $n "*) ;; *) rm -f $f;; esac
done
du -sh gpurun_out/drivers_$tag
for t in 0 1024 1040; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tuning $t > gpurun_out/bench_${tag}_t$t.json 2> gpurun_out/bench_${tag}_t$t.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}_t$t.json").read().strip().splitlines()[-1])
print("tuning $t: %.2f ms/step, accel %.2f ms, e2e %.3g, checksum %s" % (d["ms_per_step"], d["kernels"]["accel"]["ms_per_launch"], d["e2e"]["value"], d["config"]["state_checksum"]))
PY
done
ncu --set full --clock-control none --import-source on -k regex:"k_accel" -s 4 -c 2 -o gpurun_out/prof_${tag}_fm python bench.py --steps 3 --warmup 3 --max-spinup 20 --no-cpu-baseline --tuning 1024 > gpurun_out/bench_ncu_$tag.log 2>&1
ls -la gpurun_out | tail -12
