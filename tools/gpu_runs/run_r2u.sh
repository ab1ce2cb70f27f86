# multi-GPU: parity check (bit-identical forces against a single GPU) and the partitioned STMV-size bench line
OUT=gpurun_out; TAG=r2u; N=${1:-4}; shift
WL=${@:-water dhfr stmv}
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py $WL > $OUT/${TAG}_check_n$N.txt 2>&1; echo "check_n$N=$?" >> $OUT/${TAG}_status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --no-others > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench_n$N=$?" >> $OUT/${TAG}_status.txt
[ "$NOEARLY" = "1" ] && SNB_NO_EARLY_PME=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --no-others > $OUT/${TAG}_bench_n${N}_noearly.json 2> /dev/null; echo "bench_noearly_n$N=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt; grep -v "^$" $OUT/${TAG}_check_n$N.txt | tail -12; tail -3 $OUT/${TAG}_bench_n$N.err
