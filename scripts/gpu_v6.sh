# usage: bash scripts/gpu_v6.sh <tag>  -- GPU tests (timed), 10M bench, A/B of tuning switches
tag=${1:-v6}
mkdir -p gpurun_out
t0=$(date +%s)
python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -40 > gpurun_out/pytest_$tag.log
t1=$(date +%s); echo "pytest took $((t1-t0)) s"; tail -14 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
t2=$(date +%s); echo "bench took $((t2-t1)) s"; tail -2 gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json'))
print('value %.3e ms/step %.2f e2e %.3e'%(d['value'], d['ms_per_step'], d['e2e']['value']))
print('  '.join('%s=%.2f'%(k,v['ms_per_launch']) for k,v in d['kernels'].items()))
"
shift
bash scripts/gpu_ab.sh "$@"
