python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
python bench.py --no-extras > gpurun_out/bench_ld.json 2> gpurun_out/bench_ld.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_ld.json
