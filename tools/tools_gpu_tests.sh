#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu $PYTEST_ARGS > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -${TAILN:-40} gpurun_out/pytest_gpu.log
