#!/bin/bash
# A/B of library variants built into build/ab/ (gi_raytracer_b200.build.build_variant): C2 frame and a glass frame per variant, then one
# ncu --set full of k_bounce (depth 0 and 1 launches of the second warm frame) for the variants named in $NCU_VARIANTS.
# usage (on the GPU box): bash profiles/ab_variants.sh <tag> v0 v1 ...
tag=$1; shift
out=gpurun_out/ab_$tag.txt
: > $out
for v in "$@"; do
  echo "== $v caustics" >> $out
  GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1024x1024x4" >> $out
  GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene glass --spp 4 --photons 100000 >> $out 2>&1
done
for v in $NCU_VARIANTS; do
  GI_LIB=build/ab/libgi_$v.so ncu --set full --clock-control none --import-source on -k regex:"k_bounce" -s 4 -c 2 -o /tmp/full_$v -f python profiles/prof_frame.py --frames 2 > gpurun_out/ncu_${tag}_$v.log 2>&1
  { python profiles/ncu_summary.py /tmp/full_$v.ncu-rep; GI_LIB=build/ab/libgi_$v.so python profiles/sass_lines.py /tmp/full_$v.ncu-rep k_bounceILi0ELb1 40; } > gpurun_out/ncu_full_${tag}_${v}_k_bounce.txt 2>&1
done
cat $out
