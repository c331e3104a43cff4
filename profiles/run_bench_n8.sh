#!/bin/bash
# gpurun --gpus 8 --timeout 900 -- 'bash profiles/run_bench_n8.sh'  -- the headline config and BASELINE C4 (Ant, 4096 envs / 8 GPUs)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for c in hc ant; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 8 --steps 20 --warmup 5 --config $c --profile-host > gpurun_out/bench_n8_$c.json 2> gpurun_out/bench_n8_$c.err
  grep "e2e 10 rounds\|_read_stats\|end_round " gpurun_out/bench_n8_$c.err | sort | head -30
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n8_$c.json").read().strip().splitlines()[-1])
    print("$c", d["n_gpus"], "value", d["value"], d["windows_ms"], "e2e", d["e2e"]["value"], d["e2e"]["windows_ms"], d["config"]["host"])
except Exception as e:
    print("ERR", e)
PY
done
