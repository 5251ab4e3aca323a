# round 2, t20: which candidate lists go to the warp-per-query kernel (k_gather_heavy): longer than MAX_CANDS (256) unless MIN_GROUP lanes share the leaf (33 = never)
out=gpurun_out/ab_t20.txt; : > $out
for v in new mg16 mg24 mc128 mc512 mc1024; do
  if [ $v = new ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v isolated gather" >> $out; python profiles/gather_ab.py >> $out 2>&1
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 --frames 3 >> $out 2>&1
done
grep -v "^$" $out
