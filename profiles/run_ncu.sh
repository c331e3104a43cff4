#!/bin/bash
# Profiling recipe (B200_PROFILING.md) -- run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r01'
# 1) launch list of one short bench run (per-launch device time; compare SHARES, not absolutes)
# 2) --set full capture of the two kernels that matter: k_ppo_update (dominant by time) and
#    k_disc_fwdbwd (the fused discriminator kernel the north_star names)
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --cpu-rounds 0 \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ppo_update -s 2 -c 1 \
    -o gpurun_out/prof_ppo_${TAG} -f python bench.py --steps 1 --warmup 3 --cpu-rounds 0 \
    > gpurun_out/ncu_ppo_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_disc_fwdbwd -s 40 -c 2 \
    -o gpurun_out/prof_disc_${TAG} -f python bench.py --steps 1 --warmup 3 --cpu-rounds 0 \
    > gpurun_out/ncu_disc_${TAG}.log 2>&1
ls -la gpurun_out
