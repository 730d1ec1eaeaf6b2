# usage: bash scripts/gpu_v5.sh <tag>   -- GPU tests, 10M bench, ncu launch list, full ncu capture of the pair kernels
tag=${1:-v14}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -2 gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json'))
print('value %.3e ms/step %.2f e2e %.3e'%(d['value'], d['ms_per_step'], d['e2e']['value']))
print('  '.join('%s=%.2f'%(k,v['ms_per_launch']) for k,v in d['kernels'].items()))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 5 --warmup 3 --max-spinup 20 --no-cpu-baseline > gpurun_out/bench_launches_$tag.log 2>&1

ls -la gpurun_out | tail -8
