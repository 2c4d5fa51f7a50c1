#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_run49_n$N.json 2> gpurun_out/r02_run49_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r02_run49_n$N.err
python - <<PY
import json
for l in open('gpurun_out/r02_run49_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d.get('gather'), d['decode_step']['p50_us'], d['clocks'])
PY
