#!/bin/bash
# two GPUs: NCCL gather test + the N = 2 bench line (torchrun, as the driver launches it) + the reference arm under torchrun
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 600 python -m pytest tests/test_driver.py -m gpu -q -x 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_run57_n$N.json 2> gpurun_out/r02_run57_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r02_run57_n$N.err
python - <<PY
import json
for l in open('gpurun_out/r02_run57_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d.get('gather'), d['decode_step']['p50_us'], d['clocks'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-300
