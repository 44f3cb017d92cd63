#!/bin/bash
# Build a kernel-variant library for A/B timing:  tools/build_variant.sh NAME [git-rev|-] [-DPPC_...=...]
#   -> particlephysicsc_b200/lib/variants/libNAME.so (select it with PPC_LIB=...); "-" = the working tree
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; REV=${2:--}; shift; shift || true
SRC=$ROOT/particlephysicsc_b200/csrc
if [ "$REV" != "-" ]; then
  TMP=$(mktemp -d); mkdir -p $TMP/particlephysicsc_b200 $TMP/include
  git -C $ROOT archive $REV particlephysicsc_b200/csrc include | tar -x -C $TMP
  SRC=$TMP/particlephysicsc_b200/csrc
fi
mkdir -p $ROOT/particlephysicsc_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC,-Wall,-ffp-contract=off -cudart static -shared "$@" \
  -o $ROOT/particlephysicsc_b200/lib/variants/lib$NAME.so $SRC/ppc_b200.cu
echo built lib$NAME.so
