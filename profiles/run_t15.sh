# round 2, t15: gather heap with rows 0..2 in registers (new) vs all rows in shared memory (heapold)
out=gpurun_out/ab_t15.txt; : > $out
for v in heapold new heapold new; do
  if [ $v = new ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v isolated gather" >> $out; python profiles/gather_ab.py >> $out 2>&1
done
for v in heapold new; do
  if [ $v = new ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 --frames 3 >> $out 2>&1
done
unset GI_LIB
python -m pytest tests -m gpu -x -q -k "gather or golden or photon or full_size or tail or warp" > gpurun_out/gputest_t15.log 2>&1; tail -3 gpurun_out/gputest_t15.log
cat $out
