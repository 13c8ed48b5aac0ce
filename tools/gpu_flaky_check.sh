#!/bin/bash
# flakiness check: the whole GPU suite three times
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_rep$i.log 2>&1; echo "run $i rc=$?"; tail -1 gpurun_out/pytest_rep$i.log | cut -c1-150
done
