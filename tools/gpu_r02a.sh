#!/bin/bash
# r02a: GPU tests at HEAD (no -x: summaries of every free-running case are wanted), smoke, variant decomposition of the step time,
# timed-regime ncu (range replay over one round of 24 batches), launch list.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_free_running_episode_parity > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k free_running > $OUT/pytest_free.log 2>&1; echo "pytest free rc=$?" | tee -a $OUT/pytest_free.log
tail -15 $OUT/pytest_free.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
for v in "" nopair floor; do
  if [ -n "$v" ]; then export EVAC_B200_LIB=$PWD/build/variants/lib_$v.so; else unset EVAC_B200_LIB; fi
  timeout 300 python tools/step_bench.py 4096 960 24 >> $OUT/step_bench.jsonl 2>> $OUT/step_bench.err
done
unset EVAC_B200_LIB
cat $OUT/step_bench.jsonl
kill $SMI
timeout 600 ncu --replay-mode range --cache-control none --clock-control none \
  --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum \
  --csv --page raw --log-file $OUT/timed_range.csv python tools/prof_timed.py 24 1 > $OUT/timed_range.log 2>&1; echo "ncu range rc=$?"
tail -3 $OUT/timed_range.log
timeout 600 ncu --replay-mode application --cache-control none --clock-control none -k regex:evac_warp -s 96 -c 4 \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --csv --page raw --log-file $OUT/timed_app.csv python tools/prof_timed.py 24 1 > $OUT/timed_app.log 2>&1; echo "ncu app rc=$?"
tail -3 $OUT/timed_app.log
ls -la $OUT
