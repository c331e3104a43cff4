#!/bin/bash
# gpurun --timeout 600 -- 'bash profiles/run_ppo_gen_check.sh'  -- k_ppo_update_gen: parity tests, memcheck, us per step
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "ppo" 2>&1 | tail -40
timeout 300 python -m pytest tests/test_round_parity.py -q -m gpu 2>&1 | tail -30
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "ppo_update_matches_oracle and (cfg6 or cfg8)" 2>&1 | tail -15
timeout 200 python profiles/ppo_gen_timing.py > gpurun_out/ppo_gen_timing.txt 2>&1; cat gpurun_out/ppo_gen_timing.txt
