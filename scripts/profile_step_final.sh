# Round-2 evidence run on ONE B200 (gpurun): tests, the bench line of both arms, then -- only after the plain command exited
# 0 -- the ncu launch list and the whole-step `--set full` capture of the same bench command.  Numbers printed under ncu are
# never bench values.  The .ncu-rep stays in /tmp on the box (gpurun_out/ is capped at 64 MiB); its raw page, the per-kernel
# summaries and the launch list come back.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2f_tests.log
python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err || exit 1
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs"
$BENCH > gpurun_out/r2f_plain.json 2> gpurun_out/r2f_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_step_launches.csv $BENCH > gpurun_out/r2f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 130 --launch-count 100 -f -o /tmp/prof_r2 $BENCH > gpurun_out/r2f_ncu_full.log 2>&1
ncu -i /tmp/prof_r2.ncu-rep --page raw --csv > gpurun_out/r2f_step_raw.csv 2>/dev/null
python scripts/ncu_step_summary.py gpurun_out/r2f_step_raw.csv gpurun_out/r2f_step > /dev/null
python scripts/ncu_kernel_table.py gpurun_out/r2f_step_raw.csv gpurun_out/r2f_step_stalls.md "whole training step: pipes, shared memory, stall reasons"
python scripts/launch_summary.py gpurun_out/r2f_step_launches.csv > gpurun_out/r2f_step_launches_summary.txt
ls -la /tmp/prof_r2.ncu-rep gpurun_out/ | tail -20
tail -3 gpurun_out/r2f_tests.log
