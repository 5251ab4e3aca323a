# round 2, last verification of the final code: the GPU suite, then the driver's default bench line
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/r02_gputest_v6.log 2>&1; tail -3 $out/r02_gputest_v6.log
python bench.py > $out/r02v6_bench_C2.json 2> $out/r02v6_bench_C2.err; tail -c 300 $out/r02v6_bench_C2.err; cut -c1-300 $out/r02v6_bench_C2.json
