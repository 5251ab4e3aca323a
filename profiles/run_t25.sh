# round 2, t25 (second pass, with the equal-priority pair of side streams for sched_mode 0): sched_mode 2 (deferred unless the last large frame had nothing short in it): cornell must sit at the old schedule's steady time, caustics at the deferred one's
out=gpurun_out/ab_t25.txt; : > $out
python -m pytest tests/test_schedule.py -m gpu -x -q > gpurun_out/gputest_t25_sched.log 2>&1; tail -3 gpurun_out/gputest_t25_sched.log
python profiles/sched_ab.py --scenes cornell,caustics --frames 9 --variants 0:2,1:2,2:2,1:2,2:2 >> $out 2>&1
cat $out
python bench.py --config C1 --steps 5 > gpurun_out/r02v6_bench_C1.json 2> gpurun_out/r02v6_bench_C1.err; tail -c 300 gpurun_out/r02v6_bench_C1.err; cut -c1-250 gpurun_out/r02v6_bench_C1.json
