# usage (on a 2-GPU box): bash profiles/run_multi_v5.sh — final code (frame schedule of profiles/r02/README.md section 8): the 2-GPU C-ABI test and the driver's default bench line at N = 2
N=2; out=gpurun_out; P=$((29700 + N))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P"
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > $out/r02v5_test_multi_gpu.log; cat $out/r02v5_test_multi_gpu.log
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $out/r02v5_bench_C2_n$N.json 2> $out/r02v5_bench_C2_n$N.err; tail -c 300 $out/r02v5_bench_C2_n$N.err; cut -c1-330 $out/r02v5_bench_C2_n$N.json
