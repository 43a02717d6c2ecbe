#!/bin/bash
# usage: tools/ncu_capture.sh <name> <regex on the demangled kernel name> <workload> [launch-count]   (GPU box)
# Captures the matching launches with ncu --set full, exports the raw / source (SASS and CUDA-C views) /
# details pages into gpurun_out/ and drops the .ncu-rep so that gpurun_out stays under the 64 MiB cap.
set -e
name=$1; rx=$2; wl=$3; cnt=${4:-1}
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" --launch-count $cnt -f \
    -o /tmp/ncu/$name python tools/profile_run.py $wl 1 | tail -1
ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null || true
ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/${name}_details.txt
ls -la gpurun_out/${name}_*
