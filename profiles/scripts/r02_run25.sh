#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_run25_n2.json 2> gpurun_out/r02_run25_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/r02_run25_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_run25_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d.get('gather'), d['config']['parallelism'][:80])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 1 --workload dataset --clips-per-gpu 128 > gpurun_out/r02_run25_ds.json 2> gpurun_out/r02_run25_ds.err; echo "dataset rc=$?"; cut -c1-900 gpurun_out/r02_run25_ds.json
python -m pytest tests/test_driver.py -m gpu -q -x -k nccl 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
