set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in base0 rsq; do CB2_LIB=$PWD/caesar.jl_b200/variants/lib_$v.so python bench/micro/gibbs_variants.py --vars 64 > gpurun_out/var_$v.log 2>&1; done
for v in prof_base0 prof_rsq; do CB2_LIB=$PWD/caesar.jl_b200/variants/lib_$v.so timeout 300 python bench/micro/gibbs_variants.py --vars 8 --reps 1 --check 0 > gpurun_out/var_$v.log 2>&1; done
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/var_base0.log gpurun_out/var_rsq.log | tail -4
