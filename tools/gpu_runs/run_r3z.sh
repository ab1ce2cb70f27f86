# final evidence set of the round: GPU suite, full bench line, reference arm, launch lists, ncu --set full at DHFR size
OUT=gpurun_out; TAG=${1:-r3z}
mkdir -p $OUT; rm -f $OUT/${TAG}_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref=$?" >> $OUT/${TAG}_status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_dhfr.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_launch.log 2>&1; echo "launches=$?" >> $OUT/${TAG}_status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_stmv.csv \
  python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_launch_stmv.log 2>&1; echo "launches_stmv=$?" >> $OUT/${TAG}_status.txt
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_direct|k_buildList|k_pme|k_bonded|k_boundsRank|k_scanScatter|k_assignCells|k_begin|k_finish" --launch-skip 40 --launch-count 11 \
  -f -o /tmp/${TAG}_full_dhfr python bench.py --steps 3 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_dhfr.log 2>&1; echo "ncu_dhfr=$?" >> $OUT/${TAG}_status.txt
sz=$(stat -c %s /tmp/${TAG}_full_dhfr.ncu-rep 2>/dev/null || echo 99999999); if [ "$sz" -lt 40000000 ]; then cp /tmp/${TAG}_full_dhfr.ncu-rep $OUT/; fi
cat $OUT/${TAG}_status.txt; tail -4 $OUT/${TAG}_pytest.log; python tools/show_bench.py $OUT/${TAG}_bench.json | cut -c1-400
