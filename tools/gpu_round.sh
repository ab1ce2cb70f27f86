#!/bin/bash
# One GPU-box session: tests, bench, launch list and ncu --set full captures (each capture only after the same command
# exited 0 without ncu).  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh TAG [pytest] [bench] [ncu]'
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for what in "$@"; do
  case $what in
    pytest)
      timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?" >> $OUT/${TAG}_status.txt ;;
    bench)
      timeout 900 python bench.py --steps 100 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt ;;
    benchquick)
      timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-parity > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt ;;
    ref)
      timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref=$?" >> $OUT/${TAG}_status.txt ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_dhfr.csv \
        python bench.py --steps 4 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_launch.log 2>&1; echo "launches=$?" >> $OUT/${TAG}_status.txt ;;
    ncu)
      timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_direct|k_buildList|k_pme" --launch-skip 30 --launch-count 6 \
        -f -o $OUT/${TAG}_full_dhfr python bench.py --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_dhfr.log 2>&1; echo "ncu_dhfr=$?" >> $OUT/${TAG}_status.txt
      timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_directPacked|k_buildList|k_pme" --launch-skip 12 --launch-count 6 \
        -f -o $OUT/${TAG}_full_stmv python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_stmv.log 2>&1; echo "ncu_stmv=$?" >> $OUT/${TAG}_status.txt ;;
    sanitizer)
      timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_water.py > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck=$?" >> $OUT/${TAG}_status.txt
      timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_water.py > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck=$?" >> $OUT/${TAG}_status.txt ;;
  esac
done
cat $OUT/${TAG}_status.txt
