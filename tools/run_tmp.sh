#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_adhoc_scenes.py -m gpu -q -k "enclosed" -rP > gpurun_out/s2k_pytest.log 2>&1; echo "pytest exit $?"; grep -a "trace parity\|passed\|failed\|Error\|assert" gpurun_out/s2k_pytest.log | head
