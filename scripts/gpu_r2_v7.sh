# round 2, widening: five more of the reference's example drivers (KickDrift, ghosts, gradKernelCorr, Hollomon / bilinear), the ABI size guard,
# the indexed-callback variant of the e2e driver
tag=${1:-r2_v7}
mkdir -p gpurun_out/drv_$tag
WF_KEEP_DRIVER_OUT=$PWD/gpurun_out/drv_$tag timeout 1500 python -m pytest tests/test_drivers_gpu.py -m gpu -q -k "${3:-more}" 2>&1 | grep -v "^  \|^$" | cut -c1-400 | tail -120 > gpurun_out/pytest_drivers_$tag.log
tail -60 gpurun_out/pytest_drivers_$tag.log
for d in gpurun_out/drv_$tag/*/; do echo "== $d"; tail -4 $d/log.txt | cut -c1-300; done 2>/dev/null | tail -120
find gpurun_out/drv_$tag -name "*.csv" -size +100k | grep -v -E "Contact_Compression_2surf_short_full/test06_(0005|0010|0015).csv" | xargs -r rm -f      # the per-table worst entries are in report.jsonl
python -m pytest tests/test_parity_gpu.py -m gpu -q -k digest 2>&1 | tail -5
[ "$2" = "nobench" ] && exit 0
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -4 gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("default: %.2f ms/step %.3g p-steps/s; e2e %.3g (%s)" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}))
PY
