out=gpurun_out/ab_t6.txt; : > $out
for v in "$@"; do
echo "== $v caustics serial" >> $out; GI_OVERLAP_THRESHOLD=0 GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py >> $out 2>&1
echo "== $v glass serial" >> $out; GI_OVERLAP_THRESHOLD=0 GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene glass --spp 4 --photons 100000 >> $out 2>&1
done
cat $out
