#!/bin/bash
# scripts/build_variant.sh out.so [-DPMX_...=1 ...] : library with grad_umma.cu compiled under extra -D flags (A/B timing)
set -e
out=$1; shift
cd "$(dirname "$0")/.."
python -c "import proxmin_b200.build as b; b.build()" >/dev/null
C=proxmin_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas=-v "$@" -c $C/grad_umma.cu -o /tmp/grad_umma_variant.o 2>&1 | grep -A2 "k_grad_umma" | grep -i "registers\|spill" || true
objs=$(ls $C/*.o | grep -v grad_umma.o)
/usr/local/cuda/bin/nvcc -shared -o "$out" $objs /tmp/grad_umma_variant.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lpthread -lrt
