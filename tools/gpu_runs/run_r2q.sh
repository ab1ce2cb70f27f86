# A/B at DHFR / ApoA1 / FEP: per-atom PME kernels vs bricks; Estrin polynomial in the packed kernel
OUT=gpurun_out; TAG=r2q
mkdir -p $OUT
B="timeout 300 python bench.py --no-cpu --no-parity --no-others"
$B --steps 100 --warmup 5 > $OUT/${TAG}_dhfr_base.json 2>/dev/null
SNB_NO_BRICK_SPREAD=1 $B --steps 100 --warmup 5 > $OUT/${TAG}_dhfr_nobs.json 2>/dev/null
SNB_NO_BRICK_INTERP=1 $B --steps 100 --warmup 5 > $OUT/${TAG}_dhfr_nobi.json 2>/dev/null
SNB_NO_PME_BRICKS=1 $B --steps 100 --warmup 5 > $OUT/${TAG}_dhfr_nob.json 2>/dev/null
SNB_NO_PME_BRICKS=1 SNB_NO_LIST_REFINE=1 $B --steps 100 --warmup 5 > $OUT/${TAG}_dhfr_nob_noref.json 2>/dev/null
$B --workload apoa1 --steps 50 --warmup 5 > $OUT/${TAG}_apoa1_base.json 2>/dev/null
SNB_NO_BRICK_SPREAD=1 $B --workload apoa1 --steps 50 --warmup 5 > $OUT/${TAG}_apoa1_nobs.json 2>/dev/null
SNB_NO_BRICK_INTERP=1 $B --workload apoa1 --steps 50 --warmup 5 > $OUT/${TAG}_apoa1_nobi.json 2>/dev/null
SNB_NO_LIST_REFINE=1 $B --workload apoa1 --steps 50 --warmup 5 > $OUT/${TAG}_apoa1_noref.json 2>/dev/null
$B --workload fep --steps 100 --warmup 5 > $OUT/${TAG}_fep_base.json 2>/dev/null
SNB_NO_PME_BRICKS=1 $B --workload fep --steps 100 --warmup 5 > $OUT/${TAG}_fep_nob.json 2>/dev/null
SNB_NO_BRICK_SPREAD=1 $B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_nobs.json 2>/dev/null
export SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_estrin.so
$B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_estrin.json 2>/dev/null
$B --steps 100 --warmup 5 --direct-kernel packed > $OUT/${TAG}_dhfr_estrin_packed.json 2>/dev/null
unset SNB_B200_LIB
$B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_base.json 2>/dev/null
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f | head -2; done
