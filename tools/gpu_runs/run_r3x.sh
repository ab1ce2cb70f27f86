OUT=gpurun_out; TAG=r3x
mkdir -p $OUT; rm -f $OUT/${TAG}_status.txt
timeout 200 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke=$?" >> $OUT/${TAG}_status.txt
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -2 $OUT/${TAG}_smoke.log; tail -3 $OUT/${TAG}_pytest.log
