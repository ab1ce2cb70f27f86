# ncu --set full of the sort kernels at DHFR size (five kernels vs three)
OUT=gpurun_out; TAG=r3e
mkdir -p $OUT
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-others --no-parity --no-md"
SNB_FUSED_SORT=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_assignCells|k_scanScatter|k_boundsRank|k_begin|k_finish" --launch-skip 20 --launch-count 5 \
  -f -o $OUT/${TAG}_compact $B > $OUT/${TAG}_ncu1.log 2>&1; echo "ncu1=$?" >> $OUT/${TAG}_status.txt
SNB_FUSED_SORT=0 SNB_COMPACT_SORT=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_scanCells|k_scatterAtoms|k_rankInCell|k_blockBounds" --launch-skip 16 --launch-count 4 \
  -f -o $OUT/${TAG}_five $B > $OUT/${TAG}_ncu2.log 2>&1; echo "ncu2=$?" >> $OUT/${TAG}_status.txt
ls -la $OUT/${TAG}_*; cat $OUT/${TAG}_status.txt
