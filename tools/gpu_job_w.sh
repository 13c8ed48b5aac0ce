#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qnet_gpu.py -m gpu -q -x -k "tracks_the_scale" > gpurun_out/pytest_w.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_w.log | cut -c1-300
