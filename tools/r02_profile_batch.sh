set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.log 2> gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-mf --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pattern_march -s 1 -c 1 -o gpurun_out/r02_march python tools/run_march_once.py 256 > gpurun_out/ncu_march.log 2>&1
timeout 300 python tools/bench_kernels.py > gpurun_out/r02_per_kernel_timings.jsonl 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/check_march.py --small --no-oracle > gpurun_out/r02_sanitizer_memcheck_march.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/check_march.py --small --no-oracle > gpurun_out/r02_sanitizer_racecheck_march.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -x -q -m gpu -k "resize or dead_ext or tile_gather or neighbour_counts or lcn_backward" > gpurun_out/r02_sanitizer_memcheck_new_kernels.log 2>&1
timeout 600 compute-sanitizer --tool initcheck python tools/check_march.py --small --no-oracle > gpurun_out/r02_sanitizer_initcheck_march.log 2>&1
tail -3 gpurun_out/r02_sanitizer_*.log
