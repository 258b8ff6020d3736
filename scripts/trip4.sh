#!/bin/bash
# GPU tests + A/B probe
tag=${1:-trip}
mkdir -p gpurun_out
(time python -m pytest tests -q -x -m gpu) > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -6 gpurun_out/${tag}_gpu_tests.log
python scripts/ab_probe.py 2>&1 | tee gpurun_out/${tag}_ab.txt
