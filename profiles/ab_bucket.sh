# A/B runs behind DESIGN.md's kernel choices (one GPU): level-1 fan-out, grouping-kernel variants on small and large families
set -x
for b in 9 10; do
  VB_PREFILTER_L1BITS=$b python bench.py --quick --steps 6 --no-files --no-cpu-baseline > gpurun_out/b_c3_l1bits_$b.json 2>> gpurun_out/b16.err
done
for v in chain flat; do        # (the lane-refill variants 'refill' / 'sliced' of r02_ab_kernels.txt were removed after this comparison)
  VB_PREFILTER_BUCKET=$v python bench.py --workload c3_s200 --quick --steps 4 --no-files --no-cpu-baseline > gpurun_out/b_c3s200_bucket_$v.json 2>> gpurun_out/b16.err
done
python bench.py --workload c2 --quick --steps 10 --no-files --no-cpu-baseline > gpurun_out/b_c2_r16.json 2>> gpurun_out/b16.err
tail -3 gpurun_out/b16.err
