OUT=gpurun_out; TAG=r2e
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
SNB_NO_PME_BRICKS=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench_nobrick.json 2> $OUT/${TAG}_bench_nobrick.err; echo "bench_nobrick=$?" >> $OUT/${TAG}_status.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity --no-others --direct-kernel packed > $OUT/${TAG}_bench_dhfr_packed.json 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_pme|k_buildList" --launch-skip 18 --launch-count 6 \
   -f -o $OUT/${TAG}_pme_stmv python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_stmv.log 2>&1; echo "ncu_stmv=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt
