#!/bin/bash
# timing experiment on N GPUs: us per SGD step of the fused data-parallel epoch with individual exchanges disabled
# (DBMM_DP_SKIP bit mask, see p2p.cuh; results are wrong on purpose, only the timing is read)
N=${1:-2}
for m in 0 1 2 4 8 16 20 31; do
  DBMM_DP_SKIP=$m timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_time.py 2>/dev/null | tail -1 | sed "s/^/[skip=$m] /"
done
