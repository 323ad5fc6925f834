#!/bin/bash
# ncu captures of the speculative kernel on the C2 (default) or C3 frame: one full set (+ summary and per-instruction
# dump) of the second launch, and the DRAM traffic of the ninth of twelve back-to-back launches over 8 rotating buffer
# sets with the caches left alone (steady state: the write-back of earlier frames is part of the count).
#   tools/gpu_prof_spec.sh [c2|c3]
set -u
W=${1:-c2}
TASKS=187500; [ "$W" = c3 ] && TASKS=355008
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spec8 -s 1 -c 1 -f -o gpurun_out/prof_spec_$W \
  python tools/run_frames.py $W 3 > gpurun_out/prof_spec_$W.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_spec_$W.ncu-rep > gpurun_out/prof_spec_$W.txt 2>&1
python tools/ncu_sass.py gpurun_out/prof_spec_$W.ncu-rep $TASKS --dump > gpurun_out/prof_spec_${W}_sass.txt 2>&1
IPB_SETS=8 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none \
  --clock-control none -k regex:k_spec8 -s 8 -c 3 --csv --log-file gpurun_out/prof_spec_${W}_traffic.csv \
  python tools/run_frames.py $W 12 > gpurun_out/prof_spec_${W}_traffic.log 2>&1; echo "traffic rc=$?"
head -26 gpurun_out/prof_spec_$W.txt; cat gpurun_out/prof_spec_${W}_traffic.csv | tail -12
