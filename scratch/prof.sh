mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench_dhfr.json 2> gpurun_out/r1_bench_dhfr.err; echo "bench exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_dhfr.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-stmv > gpurun_out/ncu_launch.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_direct|k_buildList|k_pme|k_bonded|k_blockBounds|k_finish" --launch-skip 36 --launch-count 9 -o gpurun_out/r1_full_dhfr -f python bench.py --steps 4 --warmup 3 --no-cpu --no-stmv > gpurun_out/ncu_full.log 2>&1; echo "full exit $?"
ncu --set full --clock-control none --import-source on -k regex:"k_directPacked|k_buildList" -s 8 -c 2 -o gpurun_out/r1_full_stmv -f python bench.py --workload stmv --steps 3 --warmup 3 --no-cpu --no-stmv > gpurun_out/ncu_full_stmv.log 2>&1; echo "full stmv exit $?"
ls -la gpurun_out/*.ncu-rep
