# usage (on an N-GPU box): bash profiles/run_multi.sh N  — the 2-GPU C-ABI test (N >= 2), the driver's default bench line at N, the tile-split
# strong-scaling line of C5 at N, the NO_REDUCE A/B of the sample split, and the CLI on N GPUs against itself on one
N=$1; out=gpurun_out; P=$((29500 + N))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P"
nvidia-smi -L > $out/r02_multi_${N}_gpus.txt
if [ "$N" = "2" ]; then python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > $out/r02_test_multi_gpu.log; cat $out/r02_test_multi_gpu.log; fi
$TR bench.py --gpus $N --steps 5 --warmup 3 > $out/r02_bench_C2_n$N.json 2> $out/r02_bench_C2_n$N.err; tail -c 400 $out/r02_bench_C2_n$N.err; cut -c1-330 $out/r02_bench_C2_n$N.json
GI_BENCH_NO_REDUCE=1 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $out/r02_bench_C2_n${N}_noreduce.json 2>> $out/r02_bench_C2_n$N.err; cut -c1-200 $out/r02_bench_C2_n${N}_noreduce.json
$TR bench.py --gpus $N --config C5 --split tiles --steps 3 --warmup 3 > $out/r02_bench_C5_tiles_n$N.json 2> $out/r02_bench_C5_tiles_n$N.err; tail -c 400 $out/r02_bench_C5_tiles_n$N.err; cut -c1-330 $out/r02_bench_C5_tiles_n$N.json
$TR bench.py --gpus $N --config C2 --split tiles --steps 5 --warmup 3 > $out/r02_bench_C2_tiles_n$N.json 2> $out/r02_bench_C2_tiles_n$N.err; cut -c1-250 $out/r02_bench_C2_tiles_n$N.json
if [ "$N" = "2" ]; then
  ./gi_raytracer_b200/global-illu scenes/caustics/caustics.scn 512 512 $out/cli_1gpu.png --spp 4 --photons 200000 > $out/r02_cli.log 2>&1
  ./gi_raytracer_b200/global-illu scenes/caustics/caustics.scn 512 512 $out/cli_2gpu.png --spp 4 --photons 200000 --gpus 2 >> $out/r02_cli.log 2>&1
  cmp $out/cli_1gpu.png $out/cli_2gpu.png && echo "CLI: 2-GPU PNG identical to 1-GPU PNG" >> $out/r02_cli.log; tail -6 $out/r02_cli.log
fi
