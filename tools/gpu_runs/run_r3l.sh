OUT=gpurun_out; TAG=r3l
mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_pme_vs_ewald.py tests/test_known_answers.py tests/test_fused_sort.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest=$?" > $OUT/${TAG}_status.txt
tail -3 $OUT/${TAG}_pytest.log
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
for w in dhfr apoa1 fep; do
  $B --workload $w --steps 300 --warmup 10 > $OUT/${TAG}_${w}.json 2>$OUT/${TAG}_${w}.err
done
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,2p | cut -c1-400; done
