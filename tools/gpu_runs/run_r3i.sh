OUT=gpurun_out; TAG=r3i
mkdir -p $OUT
for cfg in "1 0" "2 1" "4 1" "4 0" "0.75 0" "0.75 1"; do
  set -- $cfg
  for w in dhfr apoa1; do echo "waves $1 priority $2" >> $OUT/${TAG}_overlap.txt; SNB_DIRECT_WAVES=$1 SNB_STREAM_PRIORITY=$2 timeout 300 python tools/overlap_probe.py $w 150 2>&1 | tail -1 | tee -a $OUT/${TAG}_overlap.txt; done
done
