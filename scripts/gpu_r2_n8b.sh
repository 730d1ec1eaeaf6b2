# round 2: BASELINE configs[3] and [4] at their stated scale on 8 GPUs: Plate-Yield_3d lattice (1e7, slabs along x) and FSW (2e7)
tag=${1:-r2_n8b}; n=${2:-8}
for w in fsw plate; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --workload $w --steps 20 --warmup 5 --max-spinup 200 > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
  tail -2 gpurun_out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$w.json").read().strip().splitlines()[-1])
    print("N=$n $w: %d particles %.3f ms/step %.4g p-steps/s, e2e %.3g; yielding %s; %s" % (d["config"]["particles"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("yielding"), d["config"]["decomposition"][:160]))
except Exception as e:
    print("bench N=$n $w failed:", e)
PY
done
