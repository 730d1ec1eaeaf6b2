set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
tail -3 gpurun_out/bench_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --max-spinup 100 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_accel|k_rate_stress|k_density<|k_thermal<|k_energy|k_nb_list" -s 12 -c 12 -o gpurun_out/prof_r1 python bench.py --steps 3 --warmup 3 --max-spinup 20 --no-cpu-baseline > gpurun_out/bench_ncu2.log 2>&1
ls -la gpurun_out
