# round 2, t21: k_direct re-reads the hit point / normal / roughness after the shadow walk instead of keeping them live across it (prev = the commit before)
out=gpurun_out/ab_t21.txt; : > $out
for v in prev new prev new; do
  if [ $v = new ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v cornell 512x512x16 depth 4" >> $out; python profiles/frame_ab.py --scene cornell --w 512 --h 512 --spp 16 --depth 4 --photons 750000 >> $out 2>&1
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 --frames 3 >> $out 2>&1
  echo "== $v foliage 1920x1080x4" >> $out; python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 4 --photons 0 --frames 3 >> $out 2>&1
  echo "== $v sponza 3840x2160x1" >> $out; python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 3 >> $out 2>&1
done
unset GI_LIB
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_t21.log 2>&1; tail -3 gpurun_out/gputest_t21.log
grep -v "^$" $out
