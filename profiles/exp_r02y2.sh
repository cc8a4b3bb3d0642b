#!/bin/bash
mkdir -p gpurun_out/drift
( time TRAJOPT_DRIFT_DUMP=gpurun_out/drift timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/r02y2_tests.log 2>&1
cat gpurun_out/r02y2_tests.log
