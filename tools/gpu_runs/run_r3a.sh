# round-2 re-entry: fused cooperative sort kernel + results stored by k_finish -- tests, then A/B on the same box
OUT=gpurun_out; TAG=r3a
mkdir -p $OUT
timeout 900 python -m pytest tests/test_fused_sort.py tests/test_device_ordered.py tests/test_overflow_retry.py tests/test_serialization.py tests/test_list_reuse.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest=$?" > $OUT/${TAG}_status.txt
tail -5 $OUT/${TAG}_pytest.log
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
$B --steps 300 --warmup 10 > $OUT/${TAG}_dhfr_base.json 2>$OUT/${TAG}_dhfr_base.err
SNB_FUSED_SORT=0 $B --steps 300 --warmup 10 > $OUT/${TAG}_dhfr_nofused.json 2>/dev/null
SNB_HOST_RESULTS=0 $B --steps 300 --warmup 10 > $OUT/${TAG}_dhfr_nohostres.json 2>/dev/null
SNB_FUSED_SORT=0 SNB_HOST_RESULTS=0 $B --steps 300 --warmup 10 > $OUT/${TAG}_dhfr_old.json 2>/dev/null
SNB_LIST_REFINE=0 $B --steps 300 --warmup 10 > $OUT/${TAG}_dhfr_norefine.json 2>/dev/null
$B --steps 300 --warmup 10 > $OUT/${TAG}_dhfr_base2.json 2>/dev/null
for w in fep apoa1; do
  $B --workload $w --steps 200 --warmup 10 > $OUT/${TAG}_${w}_base.json 2>/dev/null
  SNB_FUSED_SORT=0 SNB_HOST_RESULTS=0 $B --workload $w --steps 200 --warmup 10 > $OUT/${TAG}_${w}_old.json 2>/dev/null
done
cat $OUT/${TAG}_status.txt
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,2p | cut -c1-400; done
tail -3 $OUT/${TAG}_dhfr_base.err
