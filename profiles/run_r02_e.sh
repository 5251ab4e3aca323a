# round 2, final checkpoint (deferred frame schedule, ring of eight, shadow rays above gathers): GPU suite, the driver's bench command plain + its ncu launch list,
# the five config lines + the reference arm.  Kernel code is that of profiles/r02/ncu_full_v3_* / v4_* (unchanged).
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/r02_gputest_v5.log 2>&1; tail -3 $out/r02_gputest_v5.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/r02_bench_plain_v5.json 2> $out/r02_bench_plain_v5.err || { echo "bench failed"; tail -20 $out/r02_bench_plain_v5.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches_v5.log 2>&1
python profiles/launch_table.py $out/launches_v5.csv > $out/launches_v5.md 2>&1
python bench.py > $out/r02v5_bench_C2.json 2> $out/r02v5_bench_C2.err; tail -c 200 $out/r02v5_bench_C2.err; cut -c1-260 $out/r02v5_bench_C2.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/r02v5_bench_C2_reference.json 2>> $out/r02v5_bench_C2.err; cut -c1-200 $out/r02v5_bench_C2_reference.json
for c in C1 C3 C4 C5; do python bench.py --config $c --steps 3 > $out/r02v5_bench_$c.json 2> $out/r02v5_bench_$c.err; tail -c 200 $out/r02v5_bench_$c.err; cut -c1-250 $out/r02v5_bench_$c.json; done
