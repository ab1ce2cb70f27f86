# fused sort with the hand-written grid barrier: tests + A/B
OUT=gpurun_out; TAG=r3b
mkdir -p $OUT
timeout 900 python -m pytest tests/test_fused_sort.py tests/test_serialization.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest=$?" > $OUT/${TAG}_status.txt
tail -5 $OUT/${TAG}_pytest.log
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
for w in dhfr fep water; do
  $B --workload $w --steps 300 --warmup 10 > $OUT/${TAG}_${w}_fused.json 2>$OUT/${TAG}_${w}_fused.err
  SNB_FUSED_SORT=0 $B --workload $w --steps 300 --warmup 10 > $OUT/${TAG}_${w}_sep.json 2>/dev/null
done
cat $OUT/${TAG}_status.txt
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,2p | cut -c1-400; done
tail -3 $OUT/${TAG}_dhfr_fused.err
