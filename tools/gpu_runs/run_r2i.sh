OUT=gpurun_out; TAG=r2i
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -30 $OUT/${TAG}_pytest.log
