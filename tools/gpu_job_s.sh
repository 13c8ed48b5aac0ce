#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_augmentation.py tests/test_replay_feed.py tests/test_agent_dropin.py -m gpu -q -x > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_s.log | cut -c1-300
