#!/bin/bash
# gpurun --timeout 900 -- 'bash profiles/run_ppo_check.sh'  -- PPO kernel: whole GPU suite, phase clocks, headline bench line
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt; grep -n "^E  \|Error\|FAILED" gpurun_out/pytest_gpu.txt | head -30
IMB_VARIANT=_timing timeout 120 python profiles/ppo_phase_clocks.py > gpurun_out/ppo_phase_clocks.txt 2>&1; cat gpurun_out/ppo_phase_clocks.txt | tail -26
timeout 120 python profiles/ppo_gen_timing.py 2>&1 | head -3
timeout 300 python bench.py --config hc --steps 30 --warmup 5 --cpu-rounds 1 > gpurun_out/bench_hc.json 2> gpurun_out/bench_hc.err
tail -c 400 gpurun_out/bench_hc.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_hc.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "windows_ms")}, "e2e", d["e2e"]["value"], d["e2e"].get("windows_ms"))
print("roofline", d["roofline"]["ms_per_launch"], "stages", {k: v["ms"] for k, v in d.get("roofline_stages", {}).items()})
PY
