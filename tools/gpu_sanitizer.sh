#!/bin/bash
# compute-sanitizer over the round-2 kernels (small geometries; the 100^3 cases are excluded for time)
mkdir -p gpurun_out
SEL='f16_fp8 or (test_forward_matches_oracle_and_golden and 2) or tracks_the_scale or voxelize or depth or act_tail or fused_actor or se3 or perturb or feeder'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_qnet_gpu.py tests/test_voxelize_gpu.py tests/test_depth.py tests/test_act.py tests/test_augmentation.py tests/test_replay_feed.py -m gpu -q -x -k "$SEL and not v100 and not full_size and not graphed" > gpurun_out/sanitizer_r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_r02_memcheck.log | cut -c1-200
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_voxelize_gpu.py tests/test_act.py -m gpu -q -x -k "f16_fp8 or voxelize or act_tail" > gpurun_out/sanitizer_r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_r02_racecheck.log | cut -c1-200
