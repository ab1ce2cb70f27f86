OUT=gpurun_out; TAG=r3n
mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_pme_vs_ewald.py tests/test_slicing_invariants.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest=$?" > $OUT/${TAG}_status.txt
tail -3 $OUT/${TAG}_pytest.log
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md --steps 400 --warmup 10"
$B > $OUT/${TAG}_dhfr_fold.json 2>/dev/null
SNB_PME_FOLD=0 $B > $OUT/${TAG}_dhfr_nofold.json 2>/dev/null
SNB_LIST_REFINE=0 $B > $OUT/${TAG}_dhfr_fold_norefine.json 2>/dev/null
SNB_LIST_REFINE=0 $B --direct-kernel packed > $OUT/${TAG}_dhfr_fold_norefine_packed.json 2>/dev/null
$B > $OUT/${TAG}_dhfr_fold2.json 2>/dev/null
SNB_PME_FOLD=0 $B > $OUT/${TAG}_dhfr_nofold2.json 2>/dev/null
$B --workload fep > $OUT/${TAG}_fep_fold.json 2>/dev/null
SNB_PME_FOLD=0 $B --workload fep > $OUT/${TAG}_fep_nofold.json 2>/dev/null
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,2p | cut -c1-330; done
