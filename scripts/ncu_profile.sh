#!/bin/bash
# usage: scripts/ncu_profile.sh <tag> <bench args...>   (run under gpurun; writes gpurun_out/prof_<tag>.*)
# Captures kernels of a warmed-up bench run with --set full and exports the raw-metric and details pages as CSV.
# KREGEX picks the kernels (default: the direct engine's two kernels -- pass --engine direct; the sliced engine: KREGEX=ks_ SKIP=42 COUNT=14).
# Keep the report out of gpurun_out when it may exceed the 64 MiB that travel back: the CSV pages are what is read here anyway.
tag=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-k_graph_} -s ${SKIP:-6} -c ${COUNT:-2} -f -o gpurun_out/prof_$tag \
    python bench.py --steps 1 --warmup 3 --reads-per-step 250000 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page details --csv > gpurun_out/prof_${tag}_details.csv 2>/dev/null
