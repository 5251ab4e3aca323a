#!/bin/bash
# Round-2 profiling pass (one GPU): (1) the driver's bench command plain, then its ncu launch list; (2) ncu --set full of the hot kernels
# inside a warm C2 frame; (3) ncu --set full of the bounce kernel on C3 (glass, 41 MB of nodes + leaf records) and C5 (atrium stand-in,
# 1.2 GB) where L2 residency matters.  Only text summaries leave the box.   usage (on the GPU box): bash profiles/capture_r02.sh <tag>
tag=${1:-v1}
out=gpurun_out
mkdir -p $out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/r02_bench_plain_$tag.json 2> $out/r02_bench_plain_$tag.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$tag.log 2>&1
python profiles/launch_table.py $out/launches_$tag.csv > $out/launches_$tag.md 2>&1
python profiles/prof_frame.py --frames 2 > $out/prof_plain_$tag.log 2>&1 || exit 1
cap() {  # name regex skip count mangled-substring [prof_frame args...]
  local name=$1 rx=$2 skip=$3 cnt=$4 sym=$5; shift 5
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/full_$name -f python profiles/prof_frame.py "$@" > $out/ncu_full_${tag}_$name.log 2>&1
  { python profiles/ncu_summary.py /tmp/full_$name.ncu-rep; python profiles/sass_lines.py /tmp/full_$name.ncu-rep $sym 40; } > $out/ncu_full_${tag}_$name.txt 2>&1
  rm -f /tmp/full_$name.ncu-rep
}
cap C2_k_bounce "k_bounce" 4 2 k_bounceILi0ELb1 --frames 2
cap C2_k_direct "k_direct" 4 1 k_directILi0ELb1 --frames 2
cap C2_k_gather_sorted "k_gather_sorted" 5 1 k_gather_sorted --frames 2
cap C2_k_gather_heavy "k_gather_heavy" 5 1 k_gather_heavy --frames 2
cap C3_k_bounce "k_bounce" 0 2 k_bounceILi0ELb1 --scene glass --w 1920 --h 1080 --spp 4 --photons 275000 --frames 1
cap C5_k_bounce "k_bounce" 0 2 k_bounce_pILi0ELb1 --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 1
ls -la $out | tail -20
