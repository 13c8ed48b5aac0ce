#!/bin/bash
# ncu --set full captures of every kernel class of the forward (B=16, V=100, default math mode), one launch each:
#   gpurun --timeout 3000 -- 'bash tools/gpu_ncu_kernels.sh'; then python tools/ncu_summary.py gpurun_out/ncu_r02_<name>.ncu-rep profiles/...
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 \
    -o gpurun_out/ncu_r02_$1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-comparator > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"; ls gpurun_out/ncu_r02_$1.ncu-rep 2>/dev/null
}
cap conv3_f8c 'conv3_f8c_kernel' 1
cap upconv_gemm 'umma_gemm_kernel<.*256, .*2, .*0>' 1
cap up0_lowres 'umma_gemm_kernel<.*64, .*4, .*0>' 1
cap linear 'umma_gemm_kernel<.*256, .*2, .*1>' 40
cap geglu 'umma_gemm_kernel<.*256, .*2, .*4>' 9
cap rowmax 'umma_gemm_kernel<.*256, .*2, .*2>' 10
cap flash 'flash_attn_kernel' 3
cap patchify 'patchify_umma_kernel' 1
cap ipp 'input_preprocess_ss_kernel' 1
cap vox_scatter 'vox_scatter_kernel' 2
cap vox_fill 'vox_fill_rows_kernel' 2
cap trans_gather 'trans_gather_kernel' 1
