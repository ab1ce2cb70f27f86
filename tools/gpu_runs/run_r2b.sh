OUT=gpurun_out; TAG=r2b
export SNB_NO_PME_BRICKS=1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_nobrick.log 2>&1; echo "pytest_nobrick=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench_nobrick.json 2> $OUT/${TAG}_bench_nobrick.err; echo "bench_nobrick=$?" >> $OUT/${TAG}_status.txt
unset SNB_NO_PME_BRICKS
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity --no-others --workload dhfr > /dev/null 2>&1
cat $OUT/${TAG}_status.txt
