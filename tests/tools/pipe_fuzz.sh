ok=0; bad=0
for seed in $(seq 101 140); do
  P=$(( 40 + (seed * 37) % 900 )); MP=$(( 2 + seed % 7 )); IT=$(( 4000 + (seed * 613) % 20000 )); CH=$(( 1 + seed % 5 )); OUT=$(( 1 + seed % 11 ))
  OM=$(python -c "print([0.2,0.5,1.0,2.0,6.9][$seed % 5])"); INIT=$(( (seed / 3) % 3 )); [ $INIT -eq 1 ] && INIT=2
  spec="{\"P\": $P, \"max_par\": $MP, \"omega\": $OM, \"n_iter\": $IT, \"N\": 250, \"seed\": $seed, \"chains\": $CH, \"output\": $OUT, \"initial_network\": $INIT, \"tabulate\": $(( seed % 2 ))}"
  r=$(timeout -s KILL 120 python tests/tools/pipe_case.py "$spec" 2>&1 | tail -1)
  if [ "$r" = "OK" ]; then ok=$((ok+1)); else bad=$((bad+1)); echo "seed $seed: $r :: $spec"; fi
done
echo "two-CTA vs one-CTA, 40 random cases: $ok identical, $bad not"
