# First GPU call of the next round (one B200):   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh r2_v0'
#   1. the whole -m gpu suite (last full run: r1 v16, before the gradient-correction-with-fixed-particles, ThermalSolve and ghost-row commits;
#      those were verified test by test afterwards, never as one session)
#   2. the combinations pinned in the oracle port that have never run on a device (WF_UNVERIFIED_GPU=1)
#   3. the 10M bench line, the ncu launch list of the same command, one --set full capture of the pair kernels
tag=${1:-r2_v0}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
WF_UNVERIFIED_GPU=1 python -m pytest tests -m gpu -q -k "ghost_rows_with_gradient_correction" --tb=short 2>&1 | tail -30 > gpurun_out/pytest_unverified_$tag.log
tail -3 gpurun_out/pytest_unverified_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -2 gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --max-spinup 20 --no-cpu-baseline > gpurun_out/bench_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_accel|k_rate_stress|k_post|k_nb_list" -s 8 -c 8 -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --max-spinup 20 --no-cpu-baseline > gpurun_out/bench_ncu_$tag.log 2>&1
ls -la gpurun_out | tail -10
# 4. the packed-record micro-benchmark
nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/ubench/rec_layout.cu -o /tmp/rec_layout && /tmp/rec_layout > gpurun_out/rec_layout_$tag.txt 2>&1; cat gpurun_out/rec_layout_$tag.txt
# 5. the shared-memory-staging micro-benchmark whose numbers round 1 never committed (VERDICT "missing" 7)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/ubench/l1_gather.cu -o /tmp/l1_gather && /tmp/l1_gather > gpurun_out/l1_gather_$tag.txt 2>&1; cat gpurun_out/l1_gather_$tag.txt
