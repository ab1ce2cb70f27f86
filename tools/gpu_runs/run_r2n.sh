OUT=gpurun_out; TAG=r2n
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity --no-others --direct-kernel packed > $OUT/${TAG}_dhfr_packed.json 2> /dev/null
cat $OUT/${TAG}_status.txt; tail -5 $OUT/${TAG}_pytest.log
