# pack_weights / reduce_partials with batched loads: GPU suite, then same-box A/B against the previous build (build/prev)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in prev new prev new; do
  if [ $v = prev ]; then export GCNN_LIB_VARIANT=prev; else unset GCNN_LIB_VARIANT; fi
  python scripts/ab_option.py head_in_chain 1 1 --rounds 2 > gpurun_out/r2zg_ab_$v.json 2>> gpurun_out/r2zg_ab.err
  echo $v; python -c "import json; d=json.load(open('gpurun_out/r2zg_ab_$v.json'))['1']; print(d['step_ms'], d['combauc_forward_ms'])"
done
