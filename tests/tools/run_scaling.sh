# usage: bash tests/tools/run_scaling.sh <max_gpus> [tag]   (run under gpurun --gpus <max_gpus>)
MAXG=${1:-8}; TAG=${2:-r2}
for n in 1 2 4 8; do
  [ $n -gt $MAXG ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
  fi
  echo "n=$n rc=$?"; tail -1 gpurun_out/scale_${TAG}_n$n.json | cut -c1-250
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $MAXG --master-addr 127.0.0.1 --master-port 29700 tests/tools/gpu_dist_check.py 2>&1 | tail -2
