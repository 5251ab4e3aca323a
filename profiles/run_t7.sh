out=gpurun_out/ab_t7.txt; : > $out
for v in w0 w1 w2 w3; do
echo "== $v caustics" >> $out; GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py >> $out 2>&1
echo "== $v glass" >> $out; GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene glass --spp 4 --photons 100000 >> $out 2>&1
done
for v in w0 w2; do for mb in 0 1; do
echo "== $v mailbox $mb foliage 1920x1080x8" >> $out; GI_MAILBOX_MODE=$mb GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 8 --photons 0 --frames 2 >> $out 2>&1
echo "== $v mailbox $mb sponza 3840x2160x2" >> $out; GI_MAILBOX_MODE=$mb GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 2 --photons 0 --frames 2 >> $out 2>&1
done; done
cat $out
