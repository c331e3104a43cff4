#!/bin/bash
# gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_n2_checks.sh'
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 250 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['windows_ms'])"
timeout 120 python profiles/learning_curve.py --rounds 600 2>&1 | tail -2
timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 \
    profiles/learning_curve.py --rounds 600 2>&1 | grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM" | tail -2
