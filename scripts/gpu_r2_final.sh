# round 2, final 1-GPU validation: full -m gpu suite (GMT golden, 2 M contact digest, pinned-temperature callback), driver tests, smoke, the default bench line
tag=${1:-r2_final}
python -m pytest tests -m gpu -q --deselect tests/test_drivers_gpu.py 2>&1 | tail -30 > gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 1200 python -m pytest tests/test_drivers_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_drivers_$tag.log
tail -3 gpurun_out/pytest_drivers_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -4 gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("default: %.2f ms/step %.3g p-steps/s; e2e %.3g (%s); roofline %.4f; cpu %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}, d["roofline"]["frac"], d["cpu_baseline"]["value"]))
PY
