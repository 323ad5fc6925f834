#!/bin/bash
# one full ncu capture of the speculative kernel on the C2 frame (+ summary and per-instruction dump)
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spec8 -s 1 -c 1 -f -o gpurun_out/prof_spec_c2 \
  python tools/run_frames.py c2 3 > gpurun_out/prof_spec_c2.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_spec_c2.ncu-rep > gpurun_out/prof_spec_c2.txt 2>&1
python tools/ncu_sass.py gpurun_out/prof_spec_c2.ncu-rep 187500 --dump > gpurun_out/prof_spec_c2_sass.txt 2>&1
tail -60 gpurun_out/prof_spec_c2.txt
