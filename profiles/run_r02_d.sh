# round 2, last checkpoint (k_direct without live state across the shadow walk): the driver's bench command plain + its ncu launch list, one ncu --set full
# of k_direct on C2, the five config lines + the reference arm.  Everything else in profiles/r02/*_v3* is unchanged code.
out=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/r02_bench_plain_v4.json 2> $out/r02_bench_plain_v4.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_v4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches_v4.log 2>&1
python profiles/launch_table.py $out/launches_v4.csv > $out/launches_v4.md 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_direct" -s 4 -c 1 -o /tmp/full_direct -f python profiles/prof_frame.py --frames 2 > $out/ncu_full_v4_C2_k_direct.log 2>&1
{ python profiles/ncu_summary.py /tmp/full_direct.ncu-rep; python profiles/sass_lines.py /tmp/full_direct.ncu-rep k_directILi0ELb1 40; } > $out/ncu_full_v4_C2_k_direct.txt 2>&1
python bench.py > $out/r02v4_bench_C2.json 2> $out/r02v4_bench_C2.err; tail -c 200 $out/r02v4_bench_C2.err; cut -c1-260 $out/r02v4_bench_C2.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/r02v4_bench_C2_reference.json 2>> $out/r02v4_bench_C2.err; cut -c1-200 $out/r02v4_bench_C2_reference.json
for c in C1 C3 C4 C5; do python bench.py --config $c --steps 3 > $out/r02v4_bench_$c.json 2> $out/r02v4_bench_$c.err; tail -c 200 $out/r02v4_bench_$c.err; cut -c1-250 $out/r02v4_bench_$c.json; done
