# round 2, t14: bounce-kernel form (1 = one ray per thread, 2 = persistent warps with refetch) after the pruning, per scene
out=gpurun_out/ab_t14.txt; : > $out
for bm in 1 2; do
  export GI_BOUNCE_MODE=$bm
  echo "== bounce_mode $bm glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 --frames 3 >> $out 2>&1
  echo "== bounce_mode $bm foliage 1920x1080x4" >> $out; python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 4 --photons 0 --frames 3 >> $out 2>&1
  echo "== bounce_mode $bm sponza 3840x2160x1" >> $out; python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 3 >> $out 2>&1
done
unset GI_BOUNCE_MODE
echo "== auto: configs" >> $out
python profiles/configs.py C5 >> $out 2>&1
echo "== isolated gather (aggregated atomics, original heap)" >> $out; python profiles/gather_ab.py >> $out 2>&1
cat $out
