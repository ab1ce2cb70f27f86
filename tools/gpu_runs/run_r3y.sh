# final build: full bench line + ncu --set full at DHFR size summarised on the box
OUT=gpurun_out; TAG=r3y
mkdir -p $OUT; rm -f $OUT/${TAG}_status.txt
timeout 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_direct|k_buildList|k_pme|k_bonded|k_boundsRank|k_scanScatter|k_assignCells|k_begin|k_finish" --launch-skip 40 --launch-count 11 \
  -f -o /tmp/${TAG}_full_dhfr python bench.py --steps 3 --warmup 3 --no-cpu --no-others --no-parity --no-md > $OUT/${TAG}_ncu_dhfr.log 2>&1; echo "ncu_dhfr=$?" >> $OUT/${TAG}_status.txt
python tools/ncu_summary.py /tmp/${TAG}_full_dhfr.ncu-rep > $OUT/${TAG}_ncu_full_dhfr.txt 2>&1
python tools/stall_report.py /tmp/${TAG}_full_dhfr.ncu-rep --top 14 > $OUT/${TAG}_stalls_dhfr.txt 2>&1
ls -la /tmp/${TAG}_full_dhfr.ncu-rep
cat $OUT/${TAG}_status.txt; python tools/show_bench.py $OUT/${TAG}_bench.json | cut -c1-300
