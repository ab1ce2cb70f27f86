OUT=gpurun_out; TAG=r2m
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_pmeSpreadWarp|k_pmeInterpBrick" --launch-skip 6 --launch-count 2 \
   -f -o /tmp/${TAG}_full_stmv python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_stmv.log 2>&1; echo "ncu_stmv=$?" >> $OUT/${TAG}_status.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_pmeSpreadWarp|k_pmeInterpBrick|k_buildList" --launch-skip 15 --launch-count 3 \
   -f -o /tmp/${TAG}_full_dhfr python bench.py --steps 3 --warmup 3 --no-cpu --no-others --no-parity > $OUT/${TAG}_ncu_dhfr.log 2>&1; echo "ncu_dhfr=$?" >> $OUT/${TAG}_status.txt
ls -la /tmp/*.ncu-rep
for f in stmv dhfr; do sz=$(stat -c %s /tmp/${TAG}_full_$f.ncu-rep); if [ "$sz" -lt 25000000 ]; then cp /tmp/${TAG}_full_$f.ncu-rep $OUT/; fi; done
cat $OUT/${TAG}_status.txt
