# round 2, 1-GPU call after the default flip (contracted arithmetic is the default; the parity suite runs exact through conftest)
tag=${1:-r2_v3}
python -m pytest tests -m gpu -q --deselect tests/test_drivers_gpu.py 2>&1 | tail -30 > gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 1500 python -m pytest tests/test_drivers_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_drivers_$tag.log
tail -4 gpurun_out/pytest_drivers_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -6 gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("default: %.2f ms/step %.3g p-steps/s, accel %.2f rate %.2f post %.2f; e2e %.3g (%s); roofline %s; cpu %s; tir %s" % (d["ms_per_step"], d["value"], k["accel"]["ms_per_launch"], k["rate_stress"]["ms_per_launch"], k["post"]["ms_per_launch"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}, d["roofline"]["frac"], d["cpu_baseline"], d["config"]["state_checksum"].get("tir_nonzero_fraction")))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --max-spinup 20 --no-cpu-baseline --python-e2e > gpurun_out/bench_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_accel_fm|k_post_fm|k_rate_fm" -s 6 -c 6 -o gpurun_out/prof_${tag} python bench.py --steps 3 --warmup 3 --max-spinup 20 --no-cpu-baseline --python-e2e > gpurun_out/bench_ncu_$tag.log 2>&1
for w in plate fsw contact; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --max-spinup 200 > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
  tail -2 gpurun_out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$w.json").read().strip().splitlines()[-1])
    print("$w: %d particles %.2f ms/step %.3g p-steps/s e2e %.3g yielding %s nb %.1f" % (d["config"]["particles"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("yielding"), d["config"]["neighbours_avg"]))
except Exception as e:
    print("$w failed", e)
PY
done
ls gpurun_out | tail -5
