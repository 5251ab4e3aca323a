# round 2, t19: tail-megakernel threshold after the pruned walk (queues shorter than this finish in k_tail)
out=gpurun_out/ab_t19.txt; : > $out
for tt in 4096 8192 16384 32768 65536 131072; do
  export GI_TAIL_THRESHOLD=$tt
  echo "== tail_threshold $tt caustics 1024x1024x8" >> $out; python profiles/frame_ab.py >> $out 2>&1
  echo "== tail_threshold $tt glass 1920x1080x8" >> $out; python profiles/frame_ab.py --scene glass --w 1920 --h 1080 --spp 8 --photons 275000 --frames 3 >> $out 2>&1
done
grep -v "^$" $out
