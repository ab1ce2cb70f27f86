OUT=gpurun_out; TAG=r2j
mkdir -p $OUT
timeout 900 python -m pytest tests/test_golden_large.py tests/test_parity_gpu.py tests/test_golden.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_dhfr.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_launch.log 2>&1; echo "launches=$?" >> $OUT/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_stmv.csv \
    python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_launch_stmv.log 2>&1; echo "launches_stmv=$?" >> $OUT/${TAG}_status.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_directPacked|k_buildList|k_pmeSpreadBrick|k_pmeInterpBrick" --launch-skip 12 --launch-count 4 \
   -f -o /tmp/${TAG}_full_stmv python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_stmv.log 2>&1; echo "ncu_stmv=$?" >> $OUT/${TAG}_status.txt
ncu -i /tmp/${TAG}_full_stmv.ncu-rep --page raw --csv > $OUT/${TAG}_stmv_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_full_stmv.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/${TAG}_stmv_source.csv.gz
ls -la /tmp/${TAG}_full_stmv.ncu-rep $OUT
sz=$(stat -c %s /tmp/${TAG}_full_stmv.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/${TAG}_full_stmv.ncu-rep $OUT/; fi
cat $OUT/${TAG}_status.txt; tail -5 $OUT/${TAG}_pytest.log
