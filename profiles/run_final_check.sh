#!/bin/bash
# gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_final_check.sh'  -- the whole GPU suite incl. the 2-rank NCCL tests,
# smoke(), the headline bench line, the preference-comparison bench at 1 and 2 GPUs (members over GPUs)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt; grep -n "^E  \|Error\|FAILED\|SKIP" gpurun_out/pytest_gpu.txt | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
timeout 200 python bench.py --config pref --steps 5 --warmup 3 > gpurun_out/bench_pref_n1.json 2> gpurun_out/bench_pref.err
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --config pref --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_pref_n2.json 2>> gpurun_out/bench_pref.err
fi
tail -c 400 gpurun_out/bench_pref.err
python - <<PY
import json, os
for f in ("bench_default", "bench_pref_n1", "bench_pref_n2"):
    p = f"gpurun_out/{f}.json"
    if not os.path.exists(p):
        continue
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "steps", "warmup", "scaling")}, "e2e", d["e2e"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
