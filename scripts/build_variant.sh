#!/bin/bash
# usage: build_variant.sh name unit "flags" [unit "flags"]...
set -e
cd /root/repo/bayes_od_rc_b200
name=$1; shift
OBJ=lib/obj; V=lib/variants/obj_$name; mkdir -p $V
objs=""
declare -A repl
while [ $# -gt 0 ]; do
  unit=$1; flags=$2; shift 2
  extra=""; case $unit in k1_moments|bod_io|kp_pdq) extra="";; *) extra="-fmad=false";; esac
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fno-fast-math --ftz=false --prec-div=true --prec-sqrt=true $extra $flags -c csrc/$unit.cu -o $V/$unit.o
  repl[$unit]=1
done
for o in $OBJ/*.o; do b=$(basename $o .o); if [ -n "${repl[$b]}" ]; then objs="$objs $V/$b.o"; else objs="$objs $o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/variants/lib_$name.so $objs -cudart static
ls -la lib/variants/lib_$name.so
