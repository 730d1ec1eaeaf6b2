# round 2, final 1-GPU validation exactly as the driver runs it: the whole -m gpu suite in one invocation, smoke(), the default bench line, the reference arm
tag=${1:-r2_final2}
( time python -m pytest tests/ -x -q -m gpu ) 2>&1 | tail -12 > gpurun_out/pytest_$tag.log
tail -8 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -3 gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err
tail -1 gpurun_out/bench_${tag}_ref.json | cut -c1-600
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("default: %.2f ms/step %.3g p-steps/s; e2e %.3g (%s); roofline %.4f; cpu %.3g; launches %s; clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}, d["roofline"]["frac"], d["cpu_baseline"]["value"], d.get("gpu_launches"), d.get("clocks")))
PY
