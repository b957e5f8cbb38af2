python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -v "^    " | tail -6 > gpurun_out/final_tests.log
bash profiles/capture.sh c3 > gpurun_out/cap_c3.log 2>&1
bash profiles/capture.sh c2 > gpurun_out/cap_c2.log 2>&1
cp profiles/r02_launches_c3.txt profiles/r02_launches_c2.txt profiles/r02_kernels_ncu_full_c3.txt profiles/r02_kernels_ncu_full_c2.txt profiles/r02_kernel_lines_c3.txt profiles/r02_kernel_lines_c2.txt gpurun_out/
python bench.py > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench.err
python bench.py --workload c2 --no-files > gpurun_out/final_bench_c2.json 2>> gpurun_out/final_bench.err
tail -4 gpurun_out/final_tests.log; tail -3 gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_bench_c3.json
