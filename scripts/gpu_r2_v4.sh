# round 2, validation after: pair-uniform classes + fixed particles on the FAST path, global surface check, parallel-first-touch particle arena
tag=${1:-r2_v4}
python -m pytest tests -m gpu -q --deselect tests/test_drivers_gpu.py 2>&1 | tail -30 > gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 1200 python -m pytest tests/test_drivers_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_drivers_$tag.log
tail -4 gpurun_out/pytest_drivers_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -5 gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("default: %.2f ms/step %.3g p-steps/s, accel %.2f rate %.2f post %.2f; e2e %.3g (%s); roofline %s fp64 %s" % (d["ms_per_step"], d["value"], k["accel"]["ms_per_launch"], k["rate_stress"]["ms_per_launch"], k["post"]["ms_per_launch"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}, d["roofline"]["frac"], d["roofline"].get("fp64")))
PY
for w in fsw; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --max-spinup 200 > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
  tail -2 gpurun_out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$w.json").read().strip().splitlines()[-1])
    print("$w: %d particles %.2f ms/step %.3g p-steps/s e2e %.3g yielding %s; kernels %s" % (d["config"]["particles"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("yielding"), {n: round(v["ms_per_launch"], 2) for n, v in d["kernels"].items() if v["ms_per_launch"] > 0.5}))
except Exception as e:
    print("$w failed", e)
PY
done
