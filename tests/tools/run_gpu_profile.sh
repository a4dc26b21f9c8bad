# usage: bash tests/tools/run_gpu_profile.sh <tag> [chain|gram|sweep|launches ...]
TAG=$1; shift
CMD="python bench.py --steps 1 --warmup 1 --iters 20000 --no-cpu-baseline"
for what in "$@"; do
  case $what in
    launches)
      $CMD > gpurun_out/plain_$what.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(center|chain|column|diag|gram|graph_counts|sweep|score)" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$what.log 2>&1 ;;
    chain|gram_dmma|sweep)
      $CMD > gpurun_out/plain_$what.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${what}_kernel -c 1 -o gpurun_out/prof_${what}_$TAG $CMD > gpurun_out/ncu_$what.log 2>&1 ;;
  esac
  echo "$what rc=$?"
done
