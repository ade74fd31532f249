python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_A.json
