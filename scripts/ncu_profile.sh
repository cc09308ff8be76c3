#!/bin/bash
# usage: scripts/ncu_profile.sh <tag> <bench args...>   (run under gpurun; writes gpurun_out/prof_<tag>.*)
# Captures one k_graph_insert and one k_graph_count launch of a warmed-up bench run with --set full and exports the
# raw-metric page as CSV next to the report.
tag=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_graph_ -s 6 -c 2 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 1 --warmup 3 --reads-per-step 250000 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page details --csv > gpurun_out/prof_${tag}_details.csv 2>/dev/null
