# round 2, t18: threads per block of the thread-per-ray traversal kernels at the same 64 registers / 1024 threads per SM: 128 (default), 64, 32, 256
out=gpurun_out/ab_t18.txt; : > $out
for v in new b64 b32 b256; do
  if [ $v = new ]; then unset GI_LIB; else export GI_LIB=build/ab/libgi_$v.so; fi
  echo "== $v cornell 512x512x16 depth 4" >> $out; python profiles/frame_ab.py --scene cornell --w 512 --h 512 --spp 16 --depth 4 --photons 750000 >> $out 2>&1
  echo "== $v caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== $v glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 --frames 3 >> $out 2>&1
  echo "== $v foliage 1920x1080x4" >> $out; python profiles/frame_ab.py --scene foliage --w 1920 --h 1080 --spp 4 --photons 0 --frames 3 >> $out 2>&1
  echo "== $v sponza 3840x2160x1" >> $out; python profiles/frame_ab.py --scene sponza --w 3840 --h 2160 --spp 1 --photons 0 --frames 3 >> $out 2>&1
done
grep -v "^$" $out
