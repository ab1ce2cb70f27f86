OUT=gpurun_out; TAG=r3h
mkdir -p $OUT
for m in 0 1 2; do
  for w in dhfr apoa1 fep; do echo "priority mode $m" >> $OUT/${TAG}_overlap.txt; SNB_STREAM_PRIORITY=$m timeout 300 python tools/overlap_probe.py $w 200 2>&1 | tail -1 | tee -a $OUT/${TAG}_overlap.txt; done
done
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
for m in 0 1 2; do
  SNB_STREAM_PRIORITY=$m $B --steps 300 --warmup 10 > $OUT/${TAG}_dhfr_p$m.json 2>/dev/null
  SNB_STREAM_PRIORITY=$m $B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_p$m.json 2>/dev/null
done
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,1p | cut -c1-200; done
