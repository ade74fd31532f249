timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 python bench/micro/gibbs_variants.py --vars 256 --reps 1 --check 0 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 1 -c 1 -o gpurun_out/r02_gibbs_final python bench/micro/gibbs_variants.py --vars 256 --reps 1 --check 0 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
