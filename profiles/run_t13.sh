# round 2, t13: warp-aggregated binning atomics (new default vs previous commit's library), gather heap variants
out=gpurun_out/ab_t13.txt; : > $out
for v in old new; do
  if [ $v = new ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 >> $out 2>&1
  echo "== $v foliage 1920x1080x4" >> $out; python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 4 --photons 0 --frames 3 >> $out 2>&1
  echo "== $v sponza 3840x2160x1" >> $out; python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 3 >> $out 2>&1
done
for v in old heapold pairs floyd new; do
  if [ $v = new ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v isolated gather" >> $out; python profiles/gather_ab.py >> $out 2>&1
done
for v in heapold pairs floyd; do
  export GI_LIB=build/ab/libgi_$v.so
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
done
unset GI_LIB
python -m pytest tests -m gpu -x -q -k "gather or golden or photon or full_size or binning" > gpurun_out/gputest_t13.log 2>&1; tail -3 gpurun_out/gputest_t13.log
cat $out
