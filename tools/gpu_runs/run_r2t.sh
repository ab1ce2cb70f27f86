# whole GPU suite on the build with list reuse + ordered device entry; launch lists (insurance copy)
OUT=gpurun_out; TAG=r2t
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_dhfr.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_launch.log 2>&1; echo "launches=$?" >> $OUT/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_stmv.csv \
  python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_launch_stmv.log 2>&1; echo "launches_stmv=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -15 $OUT/${TAG}_pytest.log
