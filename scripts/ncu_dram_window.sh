#!/bin/bash
# DRAM bytes of a window of lockstep iterations of the batched engine (ncu, no replay of the skipped launches).
# usage: CFG=C5 scripts/ncu_dram_window.sh <nodes> <iters> <launch-skip> <launch-count> <out.csv>
ncu --launch-skip $3 --launch-count $4 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $5 python scripts/gpu_big_prof.py $1 $2
