# round 2, t22: frame schedule A/B (sched_mode 0 / 1 x stream priorities off / on), byte-identical accumulators; then the GPU suite
out=gpurun_out/ab_t22.txt; : > $out
python -m pytest tests/test_schedule.py -m gpu -x -q > gpurun_out/gputest_t22_sched.log 2>&1; tail -3 gpurun_out/gputest_t22_sched.log
python profiles/sched_ab.py --serial >> $out 2>&1
python profiles/sched_ab.py --scenes caustics --frames 9 --variants 0:1,1:1,0:1,1:1 >> $out 2>&1
cat $out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_t22.log 2>&1; tail -3 gpurun_out/gputest_t22.log
