#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): GPU tests, the bench line, the ncu launch list of one training step
# (per-launch durations + DRAM bytes, cold caches, serialised) and one `--set full` capture of every step kernel.
# Outputs land in gpurun_out/ (scratch); scripts/summarize_profiles.py turns them into the tracked files under profiles/.
set -x
TAG=${1:-r1b}
python -m pytest tests -m gpu -q -p no:warnings 2>&1 | tail -2 > gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
N=20480 EPOCHS=1 DBMM_GRAPH=0 DBMM_TAIL=serial python scripts/train_only.py > gpurun_out/${TAG}_plain.log 2>&1 &&
N=20480 EPOCHS=1 DBMM_GRAPH=0 DBMM_TAIL=serial ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/${TAG}_launches_step.csv python scripts/train_only.py > gpurun_out/${TAG}_ncu1.log 2>&1
N=20480 EPOCHS=1 DBMM_GRAPH=0 DBMM_TAIL=serial ncu --set full --import-source on --clock-control none \
    -k regex:"k_tail_w2|k_tail_w1|k_rows_train|k_reduce_stats|k_gemm1_tc|k_wgrad_tc" -s 120 -c 6 -o gpurun_out/${TAG}_step_full \
    python scripts/train_only.py > gpurun_out/${TAG}_ncu2.log 2>&1
for f in gpurun_out/${TAG}_tests.log gpurun_out/${TAG}_ncu1.log gpurun_out/${TAG}_ncu2.log; do tail -n 2 $f; done
