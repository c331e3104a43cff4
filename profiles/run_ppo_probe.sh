#!/bin/bash
# gpurun --timeout 600 -- 'bash profiles/run_ppo_probe.sh'  -- low-grid probe of the PPO launch + one ncu --set full capture of it
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 120 python profiles/ppo_low_grid_probe.py > gpurun_out/ppo_low_grid_probe.txt 2>&1; cat gpurun_out/ppo_low_grid_probe.txt | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ppo_update -s 1 -c 1 -o gpurun_out/prof_ppo_r02 -f python profiles/prof_ppo_single.py > gpurun_out/ncu_ppo_r02.log 2>&1; tail -2 gpurun_out/ncu_ppo_r02.log
