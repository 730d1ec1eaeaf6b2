# round 2, 8-GPU call: the scaling point the target is quoted on (device-resident + C++ Domain e2e with WELDFORM_GPUS=8), then the A/B of the
# end-of-step scalar gather (tuning bit 11: in stream order) and the 2-GPU re-run of the multi-GPU tests
tag=${1:-r2_n8}; n=${2:-8}
python -m pytest tests/test_multi_gpu.py -m gpu -q -k "goldens or host_mirror" 2>&1 | tail -5 > gpurun_out/pytest_mgpu_$tag.log; tail -3 gpurun_out/pytest_mgpu_$tag.log
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --steps 20 --warmup 5 $2 > gpurun_out/bench_${tag}_$1.json 2> gpurun_out/bench_${tag}_$1.err
  tail -2 gpurun_out/bench_${tag}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$1.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("N=$n $1: %.3f ms/step %.4g p-steps/s, e2e %.3g (%s); reduce %.3f halo %.3f ms/step; accel pass %.3f; checksum %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()},
          k["reduce"]["ms_per_launch"]*k["reduce"]["launches"]/d["steps"], k.get("halo",{"ms_per_launch":0,"launches":0})["ms_per_launch"]*k.get("halo",{"launches":0})["launches"]/d["steps"], k["accel"].get("ms_per_pass",0), d["config"]["state_checksum"]["x"]))
except Exception as e:
    print("bench N=$n $1 failed:", e)
PY
}
run default ""
run blocking_reduce "--tuning 2048 --python-e2e"
run exact "--tuning 1024 --python-e2e"
