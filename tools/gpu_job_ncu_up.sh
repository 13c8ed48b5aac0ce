#!/bin/bash
# launch list of the default bench command + ncu --set full of the up-conv GEMM (EPIK_PHASE variant) on the final binary
mkdir -p gpurun_out
bash tools/gpu_launch_list.sh
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:umma_gemm_kernel<.*256, .*2, .*5>" -s 1 -c 1 \
  -o gpurun_out/ncu_r02_upconv_gemm_v2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-comparator > gpurun_out/ncu_up.log 2>&1
echo "up rc=$?"; ls -la gpurun_out/ncu_r02_upconv_gemm_v2.ncu-rep
