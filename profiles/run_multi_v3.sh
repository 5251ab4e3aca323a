# usage (on an N-GPU box): bash profiles/run_multi_v3.sh N — final round-2 code: the 2-GPU C-ABI test + CLI (N = 2), the driver's default bench line at N
# (C2, sample split, weak) and the tile-split strong-scaling line of C5 at N
N=$1; out=gpurun_out; P=$((29600 + N))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P"
if [ "$N" = "2" ]; then python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > $out/r02v3_test_multi_gpu.log; cat $out/r02v3_test_multi_gpu.log; fi
$TR bench.py --gpus $N --steps 5 --warmup 3 > $out/r02v3_bench_C2_n$N.json 2> $out/r02v3_bench_C2_n$N.err; tail -c 300 $out/r02v3_bench_C2_n$N.err; cut -c1-330 $out/r02v3_bench_C2_n$N.json
$TR bench.py --gpus $N --config C5 --split tiles --steps 3 --warmup 3 > $out/r02v3_bench_C5_tiles_n$N.json 2> $out/r02v3_bench_C5_tiles_n$N.err; tail -c 300 $out/r02v3_bench_C5_tiles_n$N.err; cut -c1-330 $out/r02v3_bench_C5_tiles_n$N.json
if [ "$N" = "2" ]; then
  ./gi_raytracer_b200/global-illu scenes/caustics/caustics.scn 512 512 $out/cli_1gpu.png --spp 4 --photons 200000 > $out/r02v3_cli.log 2>&1
  ./gi_raytracer_b200/global-illu scenes/caustics/caustics.scn 512 512 $out/cli_2gpu.png --spp 4 --photons 200000 --gpus 2 >> $out/r02v3_cli.log 2>&1
  cmp $out/cli_1gpu.png $out/cli_2gpu.png && echo "CLI: 2-GPU PNG identical to 1-GPU PNG" >> $out/r02v3_cli.log; tail -4 $out/r02v3_cli.log
fi
