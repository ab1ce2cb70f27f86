# k_begin (copy + memset + conversion as one kernel), fused sort below 4096 cells: tests + A/B
OUT=gpurun_out; TAG=r3c
mkdir -p $OUT
timeout 900 python -m pytest tests/test_fused_sort.py tests/test_device_ordered.py tests/test_overflow_retry.py tests/test_list_reuse.py tests/test_parity_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest=$?" > $OUT/${TAG}_status.txt
tail -5 $OUT/${TAG}_pytest.log
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
for w in dhfr fep water apoa1; do
  $B --workload $w --steps 300 --warmup 10 > $OUT/${TAG}_${w}_new.json 2>$OUT/${TAG}_${w}_new.err
  SNB_BEGIN_KERNEL=0 $B --workload $w --steps 300 --warmup 10 > $OUT/${TAG}_${w}_nobegin.json 2>/dev/null
done
$B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_new.json 2>/dev/null
SNB_BEGIN_KERNEL=0 $B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_nobegin.json 2>/dev/null
cat $OUT/${TAG}_status.txt
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,2p | cut -c1-400; done
tail -3 $OUT/${TAG}_dhfr_new.err
