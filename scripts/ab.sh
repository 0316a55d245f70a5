for v in h; do echo "== $v"; CTUNET_B200_LIB=$PWD/ctunet_b200/ab/lib_$v.so python scripts/bench_kernels.py fprop 2>&1 | tail -7 ; done
