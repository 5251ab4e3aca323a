# round-2 checkpoint: GPU test suite, the driver's default bench line, the reference arm, and the four other configs (N = 1)
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_gputest_b.log; cat gpurun_out/r02_gputest_b.log
python bench.py > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err; tail -c 600 gpurun_out/r02_bench_C2.err; cut -c1-400 gpurun_out/r02_bench_C2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_C2_reference.json 2>> gpurun_out/r02_bench_C2.err; cut -c1-300 gpurun_out/r02_bench_C2_reference.json
for c in C1 C3 C4 C5; do python bench.py --config $c --steps 3 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; tail -c 300 gpurun_out/r02_bench_$c.err; cut -c1-300 gpurun_out/r02_bench_$c.json; done
