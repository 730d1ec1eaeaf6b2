# round 2, multi-GPU call (gpurun --gpus 2 or more): the whole multi-GPU test file (library slab decomposition, host mirror on WELDFORM_GPUS, plate / FSW /
# ThermalSolve goldens on two GPUs) with the log kept for profiles/, then bench lines at N = world (device-resident + C++ Domain e2e)
tag=${1:-r2_mgpu}; n=${2:-2}
python -m pytest tests/test_multi_gpu.py -m gpu -q -rA 2>&1 | tail -60 > gpurun_out/pytest_mgpu_$tag.log
grep -E "PASSED|FAILED|ERROR|passed|failed" gpurun_out/pytest_mgpu_$tag.log | tail -40
for t in 1024 0; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --steps 20 --warmup 5 --tuning $t > gpurun_out/bench_${tag}_n${n}_t$t.json 2> gpurun_out/bench_${tag}_n${n}_t$t.err
  tail -3 gpurun_out/bench_${tag}_n${n}_t$t.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_n${n}_t$t.json").read().strip().splitlines()[-1])
    print("N=$n tuning $t: %.2f ms/step %.3g p-steps/s, e2e %.3g (%s), checksum %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}, d["config"]["state_checksum"]))
except Exception as e:
    print("bench N=$n tuning $t failed:", e)
PY
done
