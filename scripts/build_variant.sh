#!/bin/bash
# Build an experimental variant of the library: the mode-specialised BSIM4 object compiled with extra flags,
# linked with the other objects of the regular build.  usage: build_variant.sh <tag> [nvcc flags...]
# -> xyce_b200/lib/exp/libxyce_b200_<tag>.so  (select it with XYCE_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/../xyce_b200/csrc"
tag=$1; shift
[ -n "$NO_GEN" ] || python3 ../../scripts/gen_spec.py . gen_spec > /dev/null
mkdir -p ../lib/exp
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr \
  --extended-lambda -Xcompiler -fPIC -fmad=true -DXB_ARITH=2 -DXB_SPEC=1 -DXB_HELPERS_INLINE "$@" \
  -c ${SRC_DIR:-gen_spec}/b4_kernels.cu -o ../lib/exp/a2x_$tag.o
objs=$(ls ../lib/obj/*.o | grep -v b4_kernels_a2x.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/exp/libxyce_b200_$tag.so $objs ../lib/exp/a2x_$tag.o -lcudart -ldl
echo built ../lib/exp/libxyce_b200_$tag.so
