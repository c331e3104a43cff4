#!/bin/bash
# gpurun --timeout 900 -- 'bash profiles/run_ncu_r02.sh'  -- launch list of one short bench run (per-launch device time:
# compare SHARES, not absolutes) + a per-kernel summary
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --windows 1 --cpu-rounds 0 > gpurun_out/bench_under_ncu_r02.log 2>&1
python - <<'PY'
import collections
import csv

rows = [r for r in csv.reader(open("gpurun_out/launches_r02.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][-70:]
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/launches_r02_summary.txt", "w") as f:
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"{v[0]:5d} launches {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  {k}"
        print(line)
        f.write(line + "\n")
PY
