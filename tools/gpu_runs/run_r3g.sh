OUT=gpurun_out; TAG=r3g
mkdir -p $OUT
for w in dhfr fep apoa1; do timeout 300 python tools/overlap_probe.py $w 200 2>&1 | tail -1 | tee -a $OUT/${TAG}_overlap.txt; done
SNB_COMPACT_SORT=0 SNB_FUSED_SORT=0 timeout 300 python tools/overlap_probe.py dhfr 200 2>&1 | tail -1 | tee -a $OUT/${TAG}_overlap.txt
