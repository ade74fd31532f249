timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cl.json 2> gpurun_out/bench_cl.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cl.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cl.json')); e=d['extra']
print(d['value'], d['roofline']['stage_ms_per_step'])
for k in ('N100','N1000'):
    h=e['hex_pose2_solve'][k]; print(k, h['solve_s'], h['device_resident_solve_s'], h['device_resident_equals_host_pointer_solve'], h['cpu_oracle_solve_s'])
print(e['c4_multihypo_products']['seconds_e2e'], e['c4_multihypo_graph'])
print(e['sap2_getsample_ms_per_particle'])
PY
