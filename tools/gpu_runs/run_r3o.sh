OUT=gpurun_out; TAG=r3o
mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fused_sort.py tests/test_golden.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest=$?" > $OUT/${TAG}_status.txt
tail -3 $OUT/${TAG}_pytest.log
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md --steps 400 --warmup 10"
for w in dhfr fep water apoa1; do $B --workload $w > $OUT/${TAG}_$w.json 2>/dev/null; done
$B > $OUT/${TAG}_dhfr2.json 2>/dev/null
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,1p | cut -c1-330; done
