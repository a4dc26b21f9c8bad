# usage: bash tests/tools/run_gpu_profile.sh <tag> [launches_default|chain|gram_dmma|sweep ...]
# Every capture is taken after the same command line exited 0 without ncu (B200_PROFILING.md).
TAG=$1; shift
CMD="python bench.py --steps 1 --warmup 1 --iters 20000 --no-cpu-baseline --no-configs"
DEF="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-kernels"
FILT="regex:^(center|chain|column|colsum|diag|gram|graph_counts|sweep|score)"
for what in "$@"; do
  case $what in
    launches_default)   # the launch list of the bench's own workload (100,000 iterations per chain)
      $DEF > gpurun_out/plain_$what.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$FILT" -c 400 --csv --log-file gpurun_out/launches_${TAG}_default.csv $DEF > gpurun_out/ncu_$what.log 2>&1 ;;
    chain|gram_dmma|sweep)
      $CMD > gpurun_out/plain_$what.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${what}_kernel -c 1 -f -o gpurun_out/prof_${what}_$TAG $CMD > gpurun_out/ncu_$what.log 2>&1 ;;
  esac
  echo "$what rc=$?"
done
