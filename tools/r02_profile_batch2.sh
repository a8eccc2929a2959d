# second profile batch of round 2 (after the LCN / smoothness / point-loss kernels): bash tools/r02_profile_batch2.sh on the GPU box
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_n1.log 2> gpurun_out/r02b_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference_arm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-mf --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
for k in lcn smooth; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${k}_ --launch-skip 3 -c 1 -o gpurun_out/r02b_${k} -f python tools/time_op.py $k > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/r02b_${k}.ncu-rep --page raw --csv > gpurun_out/r02b_${k}_raw.csv
done
timeout 300 ncu --set full --clock-control none -k regex:"box_weight|point_pattern" --launch-skip 33 -c 3 -o gpurun_out/r02b_point -f python tools/time_point_loss.py > /dev/null 2>&1
ncu -i gpurun_out/r02b_point.ncu-rep --page raw --csv > gpurun_out/r02b_point_raw.csv
timeout 300 python tools/bench_kernels.py > gpurun_out/r02b_per_kernel_timings.jsonl 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -x -q -m gpu -k "lcn or smooth or point_pattern or window_sizes" > gpurun_out/r02b_sanitizer_memcheck_lcn_smooth_point.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests -x -q -m gpu -k "smooth_loss_vs_oracle or lcn_vs_fp64 or lcn_backward" > gpurun_out/r02b_sanitizer_racecheck_lcn_smooth.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/check_march.py --small --no-oracle > gpurun_out/r02b_sanitizer_racecheck_march.log 2>&1
DIS_PARITY_REPORT=gpurun_out/r02b_parity_report.jsonl timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02b_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02b_sanitizer_*.log gpurun_out/r02b_pytest_gpu.log
