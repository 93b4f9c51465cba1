#!/bin/bash
# Builds a variant of the in-tree library with extra -D flags on ONE source file, for same-box A/B runs (TVC_LIB=...).
# usage: tools/build_variant.sh <name> <source.cu> -DFOO=1 [-DBAR=2 ...]      -> tools/build/libtinyvc_b200_<name>.so
set -e
R="$(cd "$(dirname "$0")/.." && pwd)"
N=$1; SRC=$2; shift 2
python -m tinyvc_b200.build >/dev/null
OBJ=$R/tinyvc_b200/csrc/build
mkdir -p $R/tools/build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr "$@" -c $R/tinyvc_b200/csrc/$SRC -o /tmp/variant_$N.o
OTHERS=$(ls $OBJ/*.o | grep -v "/${SRC%.cu}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $R/tools/build/libtinyvc_b200_$N.so /tmp/variant_$N.o $OTHERS -cudart static -Xlinker --no-undefined
echo "built tools/build/libtinyvc_b200_$N.so ($SRC $*)"
