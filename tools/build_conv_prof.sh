#!/bin/bash
# Test-only profile build of the library: conv_tcgen05.cu compiled with -DNPP_C3_PROF (cycle counters per role of
# conv3_kernel and per epilogue step), linked with the regular objects into tests/csrc/_bin/prof/libnpp_b200.so.
# tests/csrc/_bin/test_conv picks it up through LD_LIBRARY_PATH (tools/r02_scripts/r2_conv_cases.sh) and prints the
# counters; the shipped library is never built with the flag.
set -e
cd "$(dirname "$0")/.."
python -c "from npp_b200 import build; build.build_lib(); build.build_test_binaries()"
mkdir -p tests/csrc/_bin/prof
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda \
    -Xcompiler -fPIC -Xcompiler -fvisibility=default -I include -DNPP_C3_PROF \
    -c npp_b200/csrc/conv_tcgen05.cu -o tests/csrc/_bin/prof/conv_tcgen05_prof.o
objs=$(ls npp_b200/csrc/_obj/*.o | grep -v conv_tcgen05)
/usr/local/cuda/bin/nvcc -shared -o tests/csrc/_bin/prof/libnpp_b200.so $objs tests/csrc/_bin/prof/conv_tcgen05_prof.o
rm -f tests/csrc/_bin/prof/conv_tcgen05_prof.o
echo tests/csrc/_bin/prof/libnpp_b200.so
