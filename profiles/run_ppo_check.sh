#!/bin/bash
# gpurun --timeout 900 -- 'bash profiles/run_ppo_check.sh'  -- PPO kernel: parity tests, phase clocks, headline bench line
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ppo" 2>&1 | tail -8
timeout 300 python -m pytest tests/test_round_parity.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -6
IMB_VARIANT=_timing timeout 120 python profiles/ppo_phase_clocks.py > gpurun_out/ppo_phase_clocks.txt 2>&1; cat gpurun_out/ppo_phase_clocks.txt | tail -24
timeout 300 python bench.py --config hc --steps 30 --warmup 5 --cpu-rounds 1 > gpurun_out/bench_hc.json 2> gpurun_out/bench_hc.err
tail -c 400 gpurun_out/bench_hc.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_hc.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "windows_ms")}, "e2e", d["e2e"]["value"])
print("stages", d.get("roofline_stages"))
PY
