#!/bin/bash
# Verification of the final build on one GPU inside a small GPU budget, most important first; every step writes into gpurun_out/
# as it goes:  gpurun --timeout 840 -- 'bash tools/final_check.sh'
#   1. the GPU parity tests (through the C ABI), 2. smoke, 3. the default bench line, 4. one `ncu --set full` wave + counted flops per
#   config named in $CAPTURE (default: the BVH configs), each digested at once into gpurun_out/profiles_r02/
mkdir -p gpurun_out
T0=$SECONDS
timeout ${PYTEST_TIMEOUT:-450} python -m pytest tests -m gpu -x -v --durations=12 > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $? after $((SECONDS - T0)) s"; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err || tail -5 gpurun_out/r02_bench.err
echo "bench done after $((SECONDS - T0)) s"; cut -c1-400 gpurun_out/r02_bench.json
for c in ${CAPTURE:-C3 C2w C2}; do
  bash tools/profiles_capture.sh r02 $c > gpurun_out/final_capture_$c.log 2>&1
  echo "capture $c done after $((SECONDS - T0)) s"
done
