OUT=gpurun_out; TAG=r3p
mkdir -p $OUT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T tools/check_multi_gpu.py water fep dhfr > $OUT/${TAG}_check_n2.txt 2>&1; echo "check=$?" > $OUT/${TAG}_status.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest_multi=$?" >> $OUT/${TAG}_status.txt
timeout 600 $T bench.py --gpus 2 --steps 30 --warmup 5 > $OUT/${TAG}_bench_n2.json 2>$OUT/${TAG}_bench_n2.err; echo "bench=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; tail -12 $OUT/${TAG}_check_n2.txt; tail -3 $OUT/${TAG}_pytest_multi.log; python tools/show_bench.py $OUT/${TAG}_bench_n2.json | head -3 | cut -c1-300
