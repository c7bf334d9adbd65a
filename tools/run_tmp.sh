#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 400 python -m pytest tests -m gpu -x -q -rs > gpurun_out/s3c_pytest.log 2>&1; echo "pytest exit $? after $((SECONDS - T0)) s"; tail -8 gpurun_out/s3c_pytest.log
