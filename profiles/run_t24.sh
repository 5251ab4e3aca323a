# round 2, t24: the deferred schedule's knobs: ring of 4 / 6 / 8 hit lists, side-stream priorities (GI_STREAM_PRIO 1 / 2 / 3), tail threshold 8192 .. 524288
out=gpurun_out/ab_t24.txt; : > $out
python -m pytest tests/test_schedule.py -m gpu -x -q > gpurun_out/gputest_t24_sched.log 2>&1; tail -3 gpurun_out/gputest_t24_sched.log
python profiles/sched_ab.py --scenes caustics,glass,sponza,cornell --variants 1:1,1:1:8,1:1:4,1:2,1:3,1:1:6:8192,1:1:6:131072,1:1:6:524288 >> $out 2>&1
cat $out
