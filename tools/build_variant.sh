#!/bin/sh
# build_variant.sh <name> <extra nvcc flags...> : a copy of the library with ONLY the KG n=15 sweep unit (or, with
# GP_VARIANT_UNIT=pinn, the trainer unit) recompiled with the
# extra flags (kernel experiments; select it with GPTPINN_LIB=tools/variants/lib_<name>.so)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
OBJ=$ROOT/build/obj
VOBJ=$ROOT/build/obj_$name
mkdir -p $VOBJ $ROOT/tools/variants
cp $OBJ/*.o $VOBJ/
cd $ROOT/gpt-pinn_b200/csrc
if [ "${GP_VARIANT_UNIT:-sweep}" = "pinn" ]; then
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -diag-suppress 186 \
    "$@" -c pinn_train.cu -o $VOBJ/pinn_train.o
GP_VARIANT_N=""
fi
for n in ${GP_VARIANT_N-15}; do
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -diag-suppress 186 \
    -DGP_PDE=0 -DGP_N=$n "$@" -c sweep_inst.cu -o $VOBJ/sweep_0_$n.o
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/tools/variants/lib_$name.so $VOBJ/*.o -ccbin /usr/bin/g++
cuobjdump --dump-resource-usage $ROOT/tools/variants/lib_$name.so 2>/dev/null | grep -A1 "sweep_kernelILi0ELi15ELi[45]ELi8ELb1" | grep -o "REG:[0-9]* STACK:[0-9]*" | tr '\n' ' '
echo " <- $name (M=5, M=4 private-ring kernels)"
