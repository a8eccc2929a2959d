# final bench lines of round 2: bash tools/r02_final_bench.sh <n_gpus> on the GPU box
set -x
cd $GRAFT_REPO_ROOT
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_n1.log 2> gpurun_out/r02c_bench_n1.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-mf --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02c_bench_n$N.log 2> gpurun_out/r02c_bench_n$N.err
fi
tail -c 400 gpurun_out/r02c_bench_n$N.log
