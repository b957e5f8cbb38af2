#!/bin/bash
# Round-2 profile set of one workload (run on the GPU box from the repo root): profiles/capture.sh c3
#   launch list   ncu --metrics gpu__time_duration.sum (every kernel of one bench step; serialised, cold cache)
#   full capture  ncu --set full of one instance of every kernel of the step  -> per-kernel summary, stalls, hot lines
# Numbers printed by a bench run under ncu are never bench values.
w=${1:-c3}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_$w.csv \
    python bench.py --workload $w --quick --steps 2 --warmup 1 --no-files > gpurun_out/launches_r02_$w.log 2>&1
python profiles/ncu_summary.py launches gpurun_out/launches_r02_$w.csv > profiles/r02_launches_$w.txt
ncu --set full --clock-control none --import-source on \
    -k regex:"screen_kernel|collect_kernel|part_kernel|bucket_chain_kernel|dense_emit_kernel|build_ref_index_kernel|build_ref_text_kernel|parse_sched_kernel|expand_pairs_kernel" \
    -c 11 -o gpurun_out/r02_full_$w python bench.py --workload $w --quick --steps 1 --warmup 1 --no-files > gpurun_out/ncu_full_$w.log 2>&1
python profiles/ncu_summary.py kernels gpurun_out/r02_full_$w.ncu-rep > profiles/r02_kernels_ncu_full_$w.txt
{
  for k in parse_sched_kernel build_ref_index_kernel bucket_chain_kernel part_kernel collect_kernel; do
    echo "# $k: stall reasons + hottest lines"
    python profiles/ncu_stalls.py gpurun_out/r02_full_$w.ncu-rep "$k" 2>/dev/null | head -8
    python profiles/ncu_lines.py gpurun_out/r02_full_$w.ncu-rep "$k" 14 2>/dev/null
    echo
  done
} > profiles/r02_kernel_lines_$w.txt
tail -3 profiles/r02_launches_$w.txt
