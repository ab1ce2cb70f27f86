# round-2 evidence set: GPU suite, full bench line, reference arm, parity table, launch lists, ncu --set full captures
OUT=gpurun_out; TAG=${1:-r2w}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
timeout 1200 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref=$?" >> $OUT/${TAG}_status.txt
timeout 900 python tools/parity_table.py > $OUT/${TAG}_parity_table.json 2> $OUT/${TAG}_parity_table.md; echo "parity=$?" >> $OUT/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_dhfr.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_launch.log 2>&1; echo "launches=$?" >> $OUT/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_stmv.csv \
  python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_launch_stmv.log 2>&1; echo "launches_stmv=$?" >> $OUT/${TAG}_status.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_direct|k_buildList|k_pme|k_bonded|k_blockBounds|k_finish" --launch-skip 36 --launch-count 9 \
  -f -o /tmp/${TAG}_full_dhfr python bench.py --steps 3 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_dhfr.log 2>&1; echo "ncu_dhfr=$?" >> $OUT/${TAG}_status.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_directPacked|k_buildList|k_pmeInterpBrick|k_pmeSpread|k_pmeCombinePad|k_pmeConvolution" --launch-skip 12 --launch-count 6 \
  -f -o /tmp/${TAG}_full_stmv python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_stmv.log 2>&1; echo "ncu_stmv=$?" >> $OUT/${TAG}_status.txt
for f in dhfr stmv; do sz=$(stat -c %s /tmp/${TAG}_full_$f.ncu-rep 2>/dev/null || echo 99999999); if [ "$sz" -lt 30000000 ]; then cp /tmp/${TAG}_full_$f.ncu-rep $OUT/; fi; done
cat $OUT/${TAG}_status.txt; tail -4 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_parity_table.md
