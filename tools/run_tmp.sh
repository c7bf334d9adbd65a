#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err || tail -5 gpurun_out/r02_bench.err
echo "bench done after $((SECONDS - T0)) s"; cut -c1-300 gpurun_out/r02_bench.json
timeout 200 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
echo "reference done after $((SECONDS - T0)) s"; cut -c1-300 gpurun_out/r02_bench_reference.json
bash tools/profiles_capture.sh r02 C5 > gpurun_out/final_capture_C5.log 2>&1
echo "capture C5 done after $((SECONDS - T0)) s"
