#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_trace_parity.py -m gpu -q -k "starting_on_surfaces" -rP > gpurun_out/s2d_pytest.log 2>&1; echo "pytest exit $?"; grep -a "trace parity\|passed\|failed\|Error\|assert" gpurun_out/s2d_pytest.log | head -20
