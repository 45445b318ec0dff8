python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_1gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_1gpu.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
python tools/profile_build.py --scale 1.0 > gpurun_out/plain.log 2>&1 && ncu --nvtx --nvtx-include "step/" --metrics $M --clock-control none --csv --log-file gpurun_out/r02_final_launches.csv python tools/profile_build.py --scale 1.0 > gpurun_out/ncu.log 2>&1
python tools/profile_candidates_full.py > gpurun_out/plain_c.log 2>&1 && ncu --nvtx --nvtx-include "step/" --metrics $M --clock-control none --csv --log-file gpurun_out/r02_cand_final_launches.csv python tools/profile_candidates_full.py > gpurun_out/ncu_c.log 2>&1
# staged scatter (multi-GPU transport) on one GPU with simulated ranks: pass A / pass B times, ncu of the place pass
python tools/time_staged.py --world 8 > gpurun_out/r02_time_staged.json 2> gpurun_out/r02_time_staged.err; tail -1 gpurun_out/r02_time_staged.json
ncu --set full --clock-control none --import-source on -k regex:place_kernel -c 1 -o gpurun_out/r02_place python tools/time_staged.py --world 2 --steps 1 > gpurun_out/ncu_place.log 2>&1
echo done
