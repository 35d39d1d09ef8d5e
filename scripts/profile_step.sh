# ncu launch list + whole-step `--set full` capture of the bench step.  Run on a GPU box after `python bench.py` exited 0
# without ncu (numbers printed under ncu are never bench values).  The .ncu-rep stays in /tmp on the box (gpurun_out/ is
# capped at 64 MiB); its raw page and the per-launch list come back.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/s3_plain.json 2> gpurun_out/s3_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s3.csv $BENCH > gpurun_out/ncu_s3_list.log 2>&1
ncu --set full --clock-control none --launch-skip 130 --launch-count 90 -f -o /tmp/prof_s3 $BENCH > gpurun_out/ncu_s3_full.log 2>&1
ncu -i /tmp/prof_s3.ncu-rep --page raw --csv > gpurun_out/prof_s3_raw.csv 2>/dev/null
ls -la /tmp/prof_s3.ncu-rep gpurun_out/
