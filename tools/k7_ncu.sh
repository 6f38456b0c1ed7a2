#!/bin/bash
# One `ncu --set full` capture of the resident merge kernel on the C2 frame (run under gpurun; report lands in gpurun_out/).
# usage: tools/k7_ncu.sh <tag>
tag=${1:-k7}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:merge_fast_kernel -s 1 -c 1 -f -o gpurun_out/${tag} \
    python tools/k7_probe.py vga cvx_al 2 > gpurun_out/${tag}_ncu.log 2>&1
tail -5 gpurun_out/${tag}_ncu.log
