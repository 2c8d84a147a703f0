#!/bin/bash
# timing experiment on N GPUs: us per SGD step of the data-parallel epoch under different collective settings
N=${1:-2}
for cfg in "" "DBMM_P2P=0" "DBMM_SKIP_GRAD_AR=1" "DBMM_P2P=0 DBMM_SKIP_GRAD_AR=1"; do
  env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/dp_$N.json
  python -c "
import json; d=json.load(open('/tmp/dp_$N.json')); print('[$cfg] N=$N:', round(d['us_per_sgd_step'],1), 'us/step,', round(d['value']/1e6,2), 'M emb/s')"
done
