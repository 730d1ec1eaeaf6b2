# round 2: validation of the two-stage neighbour-list build (tentative list + uniform finalize loop) and of the pinned-temperature callback test
tag=${1:-r2_v5}
python -m pytest tests -m gpu -q --deselect tests/test_drivers_gpu.py 2>&1 | tail -30 > gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 1200 python -m pytest tests/test_drivers_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_drivers_$tag.log
tail -3 gpurun_out/pytest_drivers_$tag.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --python-e2e > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("default: %.2f ms/step %.3g p-steps/s, nb_list %.2f ms per rebuild, accel %.2f; checksum %s" % (d["ms_per_step"], d["value"], k["nb_list"]["ms_per_launch"], k["accel"]["ms_per_launch"], d["config"]["state_checksum"]["x"]))
PY
ncu --set full --clock-control none --import-source on -k regex:"k_nb_list" -c 2 -o gpurun_out/prof_${tag}_nb python bench.py --steps 3 --warmup 3 --max-spinup 20 --no-cpu-baseline --python-e2e > gpurun_out/bench_ncu_$tag.log 2>&1
ls -la gpurun_out | tail -3
