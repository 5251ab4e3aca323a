# round 2, t10: the pruned closest-hit walk (rules R1-R3 of gi_device.cuh) against the reference's full walk (-DGI_NO_PRUNE), same box
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_t10.log 2>&1; tail -5 gpurun_out/gputest_t10.log
out=gpurun_out/ab_t10.txt; : > $out
for v in noprune prune; do
  if [ $v = prune ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 >> $out 2>&1
  echo "== $v foliage 1920x1080x4" >> $out; python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 4 --photons 0 --frames 3 >> $out 2>&1
  echo "== $v sponza 3840x2160x1" >> $out; python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 3 >> $out 2>&1
done
cat $out
