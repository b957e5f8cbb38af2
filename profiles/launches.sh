#!/bin/bash
# per-launch kernel times of one bench step (ncu, serialised): profiles/launches.sh <tag> [env...]
tag=$1; shift
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --quick --steps 2 --warmup 1 > gpurun_out/launches_$tag.log 2>&1
python profiles/ncu_summary.py launches gpurun_out/launches_$tag.csv > gpurun_out/launches_$tag.txt
cat gpurun_out/launches_$tag.txt
