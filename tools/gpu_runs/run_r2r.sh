# list reuse: new tests, the whole GPU suite, bench with the md sub-object
OUT=gpurun_out; TAG=r2r
mkdir -p $OUT
timeout 600 python -m pytest tests/test_list_reuse.py -x -q > $OUT/${TAG}_pytest_reuse.log 2>&1; echo "reuse=$?" >> $OUT/${TAG}_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-parity --no-others > $OUT/${TAG}_dhfr.json 2> $OUT/${TAG}_dhfr.err; echo "dhfr=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --workload stmv --steps 30 --warmup 5 --no-cpu --no-parity --no-others > $OUT/${TAG}_stmv.json 2> $OUT/${TAG}_stmv.err; echo "stmv=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --workload apoa1 --steps 100 --warmup 5 --no-cpu --no-parity --no-others > $OUT/${TAG}_apoa1.json 2> $OUT/${TAG}_apoa1.err; echo "apoa1=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -15 $OUT/${TAG}_pytest_reuse.log; tail -5 $OUT/${TAG}_pytest.log
