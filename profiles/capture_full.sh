#!/bin/bash
# ncu --set full of each hot kernel inside the second warm C2 frame; only the text summaries leave the box (the reports are ~20 MB each).
# usage (on the GPU box): bash profiles/capture_full.sh <tag>
tag=${1:-vX}
out=gpurun_out
mkdir -p $out
python profiles/prof_frame.py --frames 2 > $out/prof_plain_$tag.log 2>&1 || exit 1
cap() {  # name regex skip count mangled-substring
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/full_$1 -f python profiles/prof_frame.py --frames 2 > $out/ncu_full_${tag}_$1.log 2>&1
  { python profiles/ncu_summary.py /tmp/full_$1.ncu-rep; python profiles/sass_lines.py /tmp/full_$1.ncu-rep $5 40; } > $out/ncu_full_${tag}_$1.txt 2>&1
  rm -f /tmp/full_$1.ncu-rep
}
cap k_bounce "k_bounce" 4 2 k_bounceILi0ELb1
cap k_gather_sorted "k_gather_sorted" 5 1 k_gather_sorted
cap k_direct "k_direct" 4 1 k_directILi0ELb1
cap k_tail "k_tail<" 1 1 k_tailILi0ELb1
cap k_tail_shadow "k_tail_shadow" 1 1 k_tail_shadowILi0ELb1
ls -la $out
