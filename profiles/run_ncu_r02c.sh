#!/bin/bash
# gpurun --timeout 240 -- 'bash profiles/run_ncu_r02c.sh'  -- end of round 2: ncu --set full capture of the final k_ppo_update
# (selected metrics + source hot lines extracted on the box), then the launch list of a short bench run
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 110 ncu --set full --clock-control none --import-source on -k regex:k_ppo_update -s 1 -c 1 -o gpurun_out/prof_ppo_r02c -f \
    python profiles/prof_ppo_single.py > gpurun_out/ncu_ppo_r02c.log 2>&1
tail -2 gpurun_out/ncu_ppo_r02c.log
python profiles/extract_selected.py gpurun_out/prof_ppo_r02c.ncu-rep gpurun_out/ncu_ppo_r02c_selected.csv
python profiles/source_hot_lines.py gpurun_out/prof_ppo_r02c.ncu-rep 45 > gpurun_out/ncu_ppo_r02c_hot_lines.txt 2>&1
head -12 gpurun_out/ncu_ppo_r02c_hot_lines.txt
grep -E "gpu__time_duration.sum|smsp__issue_active.avg.pct|launch__registers|sm__warps_active.avg.pct|dram__bytes_read.sum,|dram__bytes_write.sum," gpurun_out/ncu_ppo_r02c_selected.csv | head
rm -f gpurun_out/prof_ppo_r02c.ncu-rep
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02c.csv \
    python bench.py --steps 2 --warmup 3 --windows 1 --cpu-rounds 0 > gpurun_out/bench_under_ncu_r02c.log 2>&1
python - <<'PY'
import collections
import csv

rows = [r for r in csv.reader(open("gpurun_out/launches_r02c.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][-70:]
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/launches_r02c_summary.txt", "w") as f:
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"{v[0]:5d} launches {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  {k}"
        print(line)
        f.write(line + "\n")
PY
