#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2j_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
grep -E "passed|failed|error|10 000 edits|fma fast mode|frame [0-9]" gpurun_out/r2j_pytest.log | tail -20
tail -5 gpurun_out/r2j_pytest.log
