# list-build micro-optimisations (index arithmetic, order tests) and 13 / 14 resident warps for the packed kernel
OUT=gpurun_out; TAG=r2x
mkdir -p $OUT
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
$B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_base.json 2>/dev/null
$B --steps 100 --warmup 5 > $OUT/${TAG}_dhfr_base.json 2>/dev/null
for v in p13 p14; do
  SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_$v.so $B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_$v.json 2>/dev/null
done
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_golden_large.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -3 $OUT/${TAG}_pytest.log
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 1,2p | cut -c1-330; done
