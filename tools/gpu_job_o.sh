#!/bin/bash
# ncu --set full captures of the GEMM-engine kernel classes (B=16, V=100, default math mode), one launch each
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 \
    -o gpurun_out/ncu_r02_$1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-comparator > gpurun_out/ncu_o_$1.log 2>&1
  echo "$1 rc=$?"; ls gpurun_out/ncu_r02_$1.ncu-rep 2>/dev/null
}
cap upconv_gemm 'umma_gemm_kernel<.*256, .*2, .*0>' 1
cap up0_lowres 'umma_gemm_kernel<.*64, .*4, .*0>' 1
cap linear 'umma_gemm_kernel<.*256, .*2, .*1>' 40
cap geglu 'umma_gemm_kernel<.*256, .*2, .*4>' 9
cap rowmax 'umma_gemm_kernel<.*256, .*2, .*2>' 10
