OUT=gpurun_out; TAG=r2k
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
v=p16
SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_$v.so timeout 300 python bench.py --workload stmv --steps 30 --warmup 5 --no-cpu --no-parity --no-others > $OUT/${TAG}_stmv_$v.json 2> $OUT/${TAG}_stmv_$v.err; echo "stmv_$v=$?" >> $OUT/${TAG}_status.txt
SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_$v.so timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity --no-others --direct-kernel packed > $OUT/${TAG}_dhfr_$v.json 2> /dev/null
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity --no-others --direct-kernel packed > $OUT/${TAG}_dhfr_p12.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_dhfr.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_launch.log 2>&1; echo "launches=$?" >> $OUT/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_stmv.csv \
    python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_launch_stmv.log 2>&1; echo "launches_stmv=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -5 $OUT/${TAG}_pytest.log
