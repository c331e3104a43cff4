#!/bin/bash
# gpurun --timeout 900 -- 'bash profiles/run_pref_check.sh'  -- whole GPU suite, preference-comparison bench (fused step vs
# the autograd path), rollout launch with the tiled layers inlined / out of line
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt; grep -n "^E  \|Error\|FAILED" gpurun_out/pytest_gpu.txt | head -30
timeout 200 python bench.py --config pref --steps 5 --warmup 3 > gpurun_out/bench_pref.json 2> gpurun_out/bench_pref.err; tail -c 300 gpurun_out/bench_pref.err
IMB_PREF_AUTOGRAD=1 timeout 200 python bench.py --config pref --steps 5 --warmup 3 > gpurun_out/bench_pref_autograd.json 2>> gpurun_out/bench_pref.err
python - <<PY
import json
for f in ("bench_pref", "bench_pref_autograd"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d["config"].get("fused_step"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 200 python profiles/rollout_timing.py > gpurun_out/rollout_timing.txt 2>&1
IMB_VARIANT=_calls timeout 200 python profiles/rollout_timing.py >> gpurun_out/rollout_timing.txt 2>&1
grep -v Warning gpurun_out/rollout_timing.txt | tail -12
