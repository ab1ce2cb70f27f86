OUT=gpurun_out; TAG=r2c
export CUDA_LAUNCH_BLOCKING=1
SNB_NO_BRICK_INTERP=1 timeout 300 python tools/debug_brick.py water > $OUT/${TAG}_spread_only.log 2>&1; echo "spread_only=$?" >> $OUT/${TAG}_status.txt
SNB_NO_BRICK_SPREAD=1 timeout 300 python tools/debug_brick.py water > $OUT/${TAG}_interp_only.log 2>&1; echo "interp_only=$?" >> $OUT/${TAG}_status.txt
SNB_NO_BRICK_INTERP=1 timeout 600 compute-sanitizer --tool memcheck python tools/debug_brick.py water > $OUT/${TAG}_spread_memcheck.log 2>&1; echo "spread_memcheck=$?" >> $OUT/${TAG}_status.txt
SNB_NO_BRICK_SPREAD=1 timeout 600 compute-sanitizer --tool memcheck python tools/debug_brick.py water > $OUT/${TAG}_interp_memcheck.log 2>&1; echo "interp_memcheck=$?" >> $OUT/${TAG}_status.txt
unset CUDA_LAUNCH_BLOCKING
export SNB_NO_PME_BRICKS=1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_buildList" --launch-skip 5 --launch-count 1 \
   -f -o $OUT/${TAG}_list_dhfr python bench.py --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_dhfr.log 2>&1; echo "ncu_dhfr=$?" >> $OUT/${TAG}_status.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_buildList" --launch-skip 3 --launch-count 1 \
   -f -o $OUT/${TAG}_list_stmv python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_stmv.log 2>&1; echo "ncu_stmv=$?" >> $OUT/${TAG}_status.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity --no-others --direct-kernel packed > $OUT/${TAG}_bench_dhfr_packed.json 2>/dev/null
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity --no-others --direct-kernel general > $OUT/${TAG}_bench_dhfr_general.json 2>/dev/null
cat $OUT/${TAG}_status.txt
