# brick spreading, second version (smaller brick, more resident warps): A/B of brick shapes at STMV / ApoA1 / DHFR size + parity
OUT=gpurun_out; TAG=r2v
mkdir -p $OUT
B="timeout 300 python bench.py --no-cpu --no-parity --no-others --no-md"
$B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_peratom.json 2>/dev/null
export SNB_BRICK_SPREAD=1 SNB_PME_BRICK_MIN_ATOMS=1000
$B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_b12.json 2>/dev/null
$B --workload apoa1 --steps 50 --warmup 3 > $OUT/${TAG}_apoa1_b12.json 2>/dev/null
$B --steps 100 --warmup 5 > $OUT/${TAG}_dhfr_b12.json 2>/dev/null
for v in sb16 sb8 sb1616; do
  SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_$v.so $B --workload stmv --steps 20 --warmup 3 > $OUT/${TAG}_stmv_$v.json 2>/dev/null
done
SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_sb8.so $B --workload apoa1 --steps 50 --warmup 3 > $OUT/${TAG}_apoa1_sb8.json 2>/dev/null
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_golden_large.py tests/test_known_answers.py -m gpu -x -q > $OUT/${TAG}_pytest_brickspread.log 2>&1; echo "pytest_brickspread=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -4 $OUT/${TAG}_pytest_brickspread.log
for f in $OUT/${TAG}_*.json; do python tools/show_bench.py $f 2>/dev/null | sed -n 2p | sed "s|^|$f |"; done
