# round 2, t11: previous commit (old) vs the new walker without (noprune) and with pruning (prune), same box
out=gpurun_out/ab_t11.txt; : > $out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv >> $out
for v in old noprune prune old prune; do
  if [ $v = prune ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 >> $out 2>&1
done
cat $out
