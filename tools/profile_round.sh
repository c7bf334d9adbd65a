#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo (run under gpurun, 1 GPU).
#   tools/profile_round.sh r01        -> gpurun_out/<tag>_*  (copy the summaries into profiles/ afterwards)
TAG=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 450 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
# the wave kernels in full (one mid-render wave: generate, extend, shade_*)
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_generate|k_extend|k_shade" -s 60 -c 6 -o gpurun_out/${TAG}_wave $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.csv
ls -la gpurun_out | grep ${TAG}
