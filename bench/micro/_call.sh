for n in 48 192 296; do echo "copy CTAs $n"; CB2_LIB=$PWD/caesar.jl_b200/variants/lib_cp$n.so timeout 200 python bench/micro/e2e_probe.py 256 2>&1 | grep "call 3"; done
