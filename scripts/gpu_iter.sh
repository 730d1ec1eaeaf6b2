# usage: bash scripts/gpu_iter.sh <tag>   -- GPU tests, 10M bench, ncu of the pair kernels
tag=${1:-iter}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -2 gpurun_out/bench_$tag.err
ncu --set full --clock-control none --import-source on -k regex:"k_accel|k_rate_stress|k_post|k_nb_list" -s 8 -c 8 -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --max-spinup 20 --no-cpu-baseline > gpurun_out/bench_ncu_$tag.log 2>&1
ls -la gpurun_out | tail -8
