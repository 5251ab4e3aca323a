out=gpurun_out/ab_t5.txt; : > $out
for ov in 0 1048576; do
echo "== r01 commit caustics overlap $ov" >> $out; (cd build/r01 && GI_OVERLAP_THRESHOLD=$ov python profiles/frame_ab.py) >> $out 2>&1
echo "== r01 commit glass overlap $ov" >> $out; (cd build/r01 && GI_OVERLAP_THRESHOLD=$ov python profiles/frame_ab.py --scene glass --spp 4 --photons 100000) >> $out 2>&1
for v in v9 v4; do
echo "== $v caustics overlap $ov" >> $out; GI_OVERLAP_THRESHOLD=$ov GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py >> $out 2>&1
echo "== $v glass overlap $ov" >> $out; GI_OVERLAP_THRESHOLD=$ov GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene glass --spp 4 --photons 100000 >> $out 2>&1
done; done
cat $out
