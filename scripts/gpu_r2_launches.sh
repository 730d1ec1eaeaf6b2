# round 2, final tree: the ncu launch list of the bench command (per-launch times are cold-cache and serialised: compare the SHARES with bench.py's kernels block)
tag=${1:-r2_final}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --max-spinup 20 --no-cpu-baseline --python-e2e > gpurun_out/bench_launches_$tag.log 2>&1
python scripts/launch_shares.py gpurun_out/launches_$tag.csv > gpurun_out/launches_${tag}_summary.txt 2>&1
head -14 gpurun_out/launches_${tag}_summary.txt
