# round 2: the N = 2 bench line with both push legs of the C++ Domain API (reference-style callback and indexed callback)
tag=${1:-r2_n2b}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
grep "e2e Domain" gpurun_out/bench_${tag}.err | tail -4
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("N=2: %.2f ms/step %.3g p-steps/s; e2e %.3g (%s)" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {m: round(v.get("ms_per_step", 0), 1) for m, v in d["e2e"].get("modes", {}).items()}))
PY
