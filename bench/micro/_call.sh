timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
