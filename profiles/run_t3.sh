out=gpurun_out/ab_t3.txt; : > $out
echo "== r01 commit caustics" >> $out; (cd build/r01 && python profiles/frame_ab.py) >> $out 2>&1
echo "== r01 commit glass" >> $out; (cd build/r01 && python profiles/frame_ab.py --scene glass --spp 4 --photons 100000) >> $out 2>&1
for v in v4 v5; do for m in 1 2; do
echo "== $v bounce_mode $m caustics" >> $out; GI_BOUNCE_MODE=$m GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py >> $out 2>&1
echo "== $v bounce_mode $m glass" >> $out; GI_BOUNCE_MODE=$m GI_LIB=build/ab/libgi_$v.so python profiles/frame_ab.py --scene glass --spp 4 --photons 100000 >> $out 2>&1
done; done
cat $out
