#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2n_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/s2n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
