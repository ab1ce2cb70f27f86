# compact sort (3 kernels), fused sort with two barriers: tests + A/B
OUT=gpurun_out; TAG=r3d
mkdir -p $OUT
timeout 900 python -m pytest tests/test_fused_sort.py tests/test_device_ordered.py tests/test_golden.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest=$?" > $OUT/${TAG}_status.txt
tail -5 $OUT/${TAG}_pytest.log
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
for w in dhfr fep water; do
  $B --workload $w --steps 300 --warmup 10 > $OUT/${TAG}_${w}_new.json 2>$OUT/${TAG}_${w}_new.err
  SNB_COMPACT_SORT=0 SNB_FUSED_SORT=0 $B --workload $w --steps 300 --warmup 10 > $OUT/${TAG}_${w}_five.json 2>/dev/null
done
SNB_FUSED_SORT=0 $B --workload fep --steps 300 --warmup 10 > $OUT/${TAG}_fep_compact.json 2>/dev/null
SNB_FUSED_SORT=0 $B --workload water --steps 300 --warmup 10 > $OUT/${TAG}_water_compact.json 2>/dev/null
cat $OUT/${TAG}_status.txt
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,2p | cut -c1-400; done
tail -3 $OUT/${TAG}_dhfr_new.err
