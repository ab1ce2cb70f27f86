OUT=gpurun_out; TAG=r3q
mkdir -p $OUT
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
$B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_early.json 2>/dev/null
SNB_EARLY_GRID_ZERO=0 $B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_late.json 2>/dev/null
SNB_EARLY_GRID_ZERO=0 $B --workload apoa1 --steps 300 --warmup 10 > $OUT/${TAG}_apoa1_late.json 2>/dev/null
$B --workload apoa1 --steps 300 --warmup 10 > $OUT/${TAG}_apoa1_early.json 2>/dev/null
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,1p | cut -c1-200; done
