# round 2, t12: one pop / load / drop site in the walker; variants: previous commit (old), full walk (noprune), R2 on leaves only, 72 / 80 registers
out=gpurun_out/ab_t12.txt; : > $out
for v in old noprune prune r2leaf minb7 minb6; do
  if [ $v = prune ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 >> $out 2>&1
  echo "== $v foliage 1920x1080x4" >> $out; python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 4 --photons 0 --frames 3 >> $out 2>&1
  echo "== $v sponza 3840x2160x1" >> $out; python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 3 >> $out 2>&1
done
unset GI_LIB
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_t12.log 2>&1; tail -3 gpurun_out/gputest_t12.log
cat $out
