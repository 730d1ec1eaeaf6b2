# round 2, widening: five more of the reference's example drivers (KickDrift, ghosts, gradKernelCorr, Hollomon / bilinear), the ABI size guard,
# the indexed-callback variant of the e2e driver
tag=${1:-r2_v7}
mkdir -p gpurun_out/drv_$tag
WF_KEEP_DRIVER_OUT=gpurun_out/drv_$tag timeout 1500 python -m pytest tests/test_drivers_gpu.py -m gpu -q -k more 2>&1 | tail -60 > gpurun_out/pytest_drivers_$tag.log
tail -25 gpurun_out/pytest_drivers_$tag.log
find gpurun_out/drv_$tag -name "test06_*.csv" | grep -v -E "_(0001|0002|0003|0004|0020|0040|0060).csv" | xargs -r rm -f
find gpurun_out/drv_$tag -name "*.csv" -size +3M -delete
python -m pytest tests/test_abi.py tests/test_host_api.py -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -4 gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("default: %.2f ms/step %.3g p-steps/s; e2e %.3g (%s)" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}))
PY
