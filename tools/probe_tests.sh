#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "ragged or peer_layout or fails_early" > gpurun_out/tests_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_new.log
tail -12 gpurun_out/tests_new.log
