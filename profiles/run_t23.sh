# round 2, t23: frame schedule A/B, second pass: sched_mode 0 (equal priorities) / 1 / 2 (main stream at the higher priority), byte-identical accumulators
out=gpurun_out/ab_t23.txt; : > $out
python -m pytest tests/test_schedule.py -m gpu -x -q > gpurun_out/gputest_t23_sched.log 2>&1; tail -3 gpurun_out/gputest_t23_sched.log
python profiles/sched_ab.py >> $out 2>&1
cat $out
