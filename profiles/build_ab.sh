#!/bin/bash
# A/B builds of libbmc_b200.so with extra -D flags (not tracked; they travel to the GPU box with the snapshot):
#   bash profiles/build_ab.sh nocb -DBMC_CONST_BANK_D=0
# then  BMC_LIB=profiles/ab/libbmc_nocb.so python profiles/time_gibbs.py ...
set -e
name=$1; shift
mkdir -p profiles/ab
cd pybmc_b200/csrc
objs=""
for u in linalg gibbs simplex predict literal; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
        -c $u.cu -o /tmp/ab_${name}_$u.o &
    objs="$objs /tmp/ab_${name}_$u.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../profiles/ab/libbmc_${name}.so $objs -cudart static
echo profiles/ab/libbmc_${name}.so
