OUT=gpurun_out; TAG=r2o; N=${1:-2}
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py water fep dhfr stmv > $OUT/${TAG}_check_n$N.txt 2>&1; echo "check_n$N=$?" >> $OUT/${TAG}_status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench_n$N=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; grep -v "^$" $OUT/${TAG}_check_n$N.txt | tail -15
