#!/bin/bash
# GPU-box: the -m gpu suite only (final check of the shipped tree)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -n 12 gpurun_out/pytest.log | cut -c1-400
