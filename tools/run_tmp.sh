#!/bin/bash
# scratch command file for short gpurun calls (edit, then: gpurun --timeout 300 -- 'bash tools/run_tmp.sh')
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rs > gpurun_out/tmp_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/tmp_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
