out=gpurun_out/ab_t8.txt; : > $out
for v in p20 p12 p26; do for m in 1 2; do
[ "$m" = "1" ] && [ "$v" != "p20" ] && continue
echo "== $v bounce_mode $m caustics" >> $out; GI_BOUNCE_MODE=$m GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py >> $out 2>&1
echo "== $v bounce_mode $m glass" >> $out; GI_BOUNCE_MODE=$m GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene glass --spp 4 --photons 100000 >> $out 2>&1
echo "== $v bounce_mode $m foliage" >> $out; GI_BOUNCE_MODE=$m GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 4 --photons 0 --frames 2 >> $out 2>&1
echo "== $v bounce_mode $m sponza" >> $out; GI_BOUNCE_MODE=$m GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 2 >> $out 2>&1
done; done
cat $out
