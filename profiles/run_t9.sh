out=gpurun_out/ab_t9.txt; : > $out
for bt in 65536 0; do
echo "== bin_threshold $bt foliage" >> $out; GI_BIN_THRESHOLD=$bt python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 4 --photons 0 --frames 3 >> $out 2>&1
echo "== bin_threshold $bt sponza" >> $out; GI_BIN_THRESHOLD=$bt python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 3 >> $out 2>&1
done
cat $out
