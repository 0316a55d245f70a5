echo "== product"; python scripts/bench_kernels.py both 2>&1 | tail -7
for v in i; do echo "== $v"; CTUNET_B200_LIB=$PWD/ctunet_b200/ab/lib_$v.so python scripts/bench_kernels.py fprop 2>&1 | tail -7 ; done
