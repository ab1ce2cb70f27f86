OUT=gpurun_out; TAG=r2f
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_golden_large.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
for v in p14 p16; do
  SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_$v.so timeout 300 python bench.py --workload stmv --steps 30 --warmup 5 --no-cpu --no-parity --no-others > $OUT/${TAG}_stmv_$v.json 2> $OUT/${TAG}_stmv_$v.err; echo "stmv_$v=$?" >> $OUT/${TAG}_status.txt
  SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_$v.so timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity --no-others --direct-kernel packed > $OUT/${TAG}_dhfr_$v.json 2> /dev/null
done
timeout 300 python bench.py --workload stmv --steps 30 --warmup 5 --no-cpu --no-parity --no-others > $OUT/${TAG}_stmv_p12.json 2> /dev/null
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_pme" --launch-skip 15 --launch-count 5 \
   -f -o $OUT/${TAG}_pme_stmv python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_stmv.log 2>&1; echo "ncu_stmv=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt
