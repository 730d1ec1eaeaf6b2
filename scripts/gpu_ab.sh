# A/B of kernel tuning switches on the 10M bench (short runs, no CPU baseline)
mkdir -p gpurun_out
for t in "$@"; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tuning $t > gpurun_out/ab_$t.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$t.json"))
print("tuning $t: %.3e p-steps/s  %.2f ms/step | "%(d["value"], d["ms_per_step"]) + "  ".join("%s=%.2f"%(k,v["ms_per_launch"]) for k,v in d["kernels"].items() if v["ms_per_launch"]>0.3))
PY
done
