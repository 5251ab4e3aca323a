# round 2, final checkpoint (pruned walk in its two loop forms, re-tuned bounce form, aggregated atomics, JPEG textures): GPU suite, profiling pass v3, the five config lines + reference arm
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_gputest_v3.log; cat gpurun_out/r02_gputest_v3.log
bash profiles/capture_r02.sh v3 > gpurun_out/capture_r02_v3.log 2>&1; tail -3 gpurun_out/capture_r02_v3.log
python bench.py > gpurun_out/r02v3_bench_C2.json 2> gpurun_out/r02v3_bench_C2.err; tail -c 300 gpurun_out/r02v3_bench_C2.err; cut -c1-300 gpurun_out/r02v3_bench_C2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02v3_bench_C2_reference.json 2>> gpurun_out/r02v3_bench_C2.err; cut -c1-300 gpurun_out/r02v3_bench_C2_reference.json
for c in C1 C3 C4 C5; do python bench.py --config $c --steps 3 > gpurun_out/r02v3_bench_$c.json 2> gpurun_out/r02v3_bench_$c.err; tail -c 200 gpurun_out/r02v3_bench_$c.err; cut -c1-250 gpurun_out/r02v3_bench_$c.json; done
