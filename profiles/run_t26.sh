# round 2, t26: k_gather_heavy on an auxiliary stream beside the counting sort + k_gather_sorted (the locate kernel queues the long lists)
out=gpurun_out/ab_t26.txt; : > $out
python -m pytest tests/test_schedule.py tests/test_gpu_parity.py -m gpu -x -q -k "schedule or gather or photon_map or golden" > gpurun_out/gputest_t26.log 2>&1; tail -3 gpurun_out/gputest_t26.log
for v in aux noaux aux noaux; do
  if [ $v = noaux ]; then export GI_NO_GATHER_AUX=1; else unset GI_NO_GATHER_AUX; fi
  echo "== $v" >> $out
  python profiles/gather_ab.py >> $out 2>&1
  python profiles/sched_ab.py --scenes caustics,glass --variants 2:2 --frames 7 >> $out 2>&1
done
unset GI_NO_GATHER_AUX
grep -v "^$" $out
