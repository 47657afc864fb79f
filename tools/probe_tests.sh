#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_dropin_manager.py -x -q -m gpu -s -k "interleaved or ten_thousand" > gpurun_out/tests_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_new.log
tail -25 gpurun_out/tests_new.log
