# round 2: two-GPU regression after making the migration / halo capacity failures collective (comm_agree): the whole multi-GPU suite, the reference drivers
# on two GPUs, and the N = 2 bench line (state checksums must equal the N = 1 line's)
tag=${1:-r2_mgpu3}
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_mgpu_$tag.log
tail -3 gpurun_out/pytest_mgpu_$tag.log
timeout 900 python -m pytest tests/test_drivers_gpu.py -m gpu -q -k two_gpus 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_${tag}_n2.json 2> gpurun_out/bench_${tag}_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}_n2.json").read().strip().splitlines()[-1])
print("N=2: %.2f ms/step %.3g p-steps/s checksum %s" % (d["ms_per_step"], d["value"], d["config"].get("state_checksum")))
PY
