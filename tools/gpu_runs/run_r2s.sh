# tests (whole GPU suite), refine A/B at DHFR size, STMV trajectory with the refined kept list, sanitizer logs, launch lists
OUT=gpurun_out; TAG=r2s
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt
B="timeout 300 python bench.py --no-cpu --no-parity --no-others"
$B --steps 100 --warmup 5 --no-md > $OUT/${TAG}_dhfr_base.json 2>/dev/null
SNB_LIST_REFINE=0 $B --steps 100 --warmup 5 --no-md > $OUT/${TAG}_dhfr_noref.json 2>/dev/null
SNB_LIST_REFINE=0 $B --workload fep --steps 100 --warmup 5 --no-md > $OUT/${TAG}_fep_noref.json 2>/dev/null
$B --workload fep --steps 100 --warmup 5 --no-md > $OUT/${TAG}_fep_base.json 2>/dev/null
$B --workload stmv --steps 30 --warmup 5 > $OUT/${TAG}_stmv.json 2>/dev/null
$B --workload stmv --steps 30 --warmup 5 --list-skin 0.06 > $OUT/${TAG}_stmv_skin06.json 2>/dev/null
$B --steps 100 --warmup 5 --list-skin 0.06 > $OUT/${TAG}_dhfr_skin06.json 2>/dev/null
timeout 600 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -5 $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_smoke.log; tail -3 $OUT/${TAG}_memcheck.log; tail -3 $OUT/${TAG}_racecheck.log
