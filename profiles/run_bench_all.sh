#!/bin/bash
# gpurun --timeout 1500 -- 'bash profiles/run_bench_all.sh'   -- GPU test suite + one short bench line per BASELINE config
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for c in hc cartpole airl_hc ant pref; do
  echo "== $c"
  timeout 300 python bench.py --config $c --steps 20 --warmup 5 --cpu-rounds 2 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -c 800 gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$c.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "windows_ms", "gpu_launches")}, "e2e", d["e2e"]["value"], "cpu",
          (d.get("cpu_baseline") or {}).get("value"))
    print("roof", (d.get("roofline") or {}).get("ms_per_launch"), "disc",
          {k: (d.get("roofline_disc") or {}).get(k) for k in ("Grows_per_s", "frac", "speedup_vs_ffma_kernel")})
    print("stages", d.get("roofline_stages"))
except Exception as e:
    print("ERR", e)
PY
done
