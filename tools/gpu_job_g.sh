#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_agent_dropin.py -m gpu -q -x -s > gpurun_out/pytest_agent.log 2>&1; echo "agent rc=$?"
tail -40 gpurun_out/pytest_agent.log | cut -c1-250
