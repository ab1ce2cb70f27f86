OUT=gpurun_out; TAG=r3m
mkdir -p $OUT
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md --steps 400 --warmup 10"
$B > $OUT/${TAG}_dhfr_base.json 2>/dev/null
SNB_LIST_REFINE=0 $B > $OUT/${TAG}_dhfr_norefine.json 2>/dev/null
$B --direct-kernel packed > $OUT/${TAG}_dhfr_packed.json 2>/dev/null
SNB_LIST_REFINE=0 $B --direct-kernel packed > $OUT/${TAG}_dhfr_packed_norefine.json 2>/dev/null
$B > $OUT/${TAG}_dhfr_base2.json 2>/dev/null
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,2p | cut -c1-330; done
