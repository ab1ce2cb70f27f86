OUT=gpurun_out; TAG=r3k
mkdir -p $OUT
T="timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k test_water_box"
$T > $OUT/${TAG}_plain.log 2>&1; echo "plain=$?" >> $OUT/${TAG}_status.txt
CUDA_LAUNCH_BLOCKING=1 SNB_NO_GRAPH=1 $T > $OUT/${TAG}_blocking.log 2>&1; echo "blocking=$?" >> $OUT/${TAG}_status.txt
SNB_FUSED_SORT=0 $T > $OUT/${TAG}_nofused.log 2>&1; echo "nofused=$?" >> $OUT/${TAG}_status.txt
SNB_FUSED_SORT=0 SNB_COMPACT_SORT=0 $T > $OUT/${TAG}_five.log 2>&1; echo "five=$?" >> $OUT/${TAG}_status.txt
SNB_BEGIN_KERNEL=0 $T > $OUT/${TAG}_nobegin.log 2>&1; echo "nobegin=$?" >> $OUT/${TAG}_status.txt
SNB_HOST_RESULTS=0 $T > $OUT/${TAG}_nohostres.log 2>&1; echo "nohostres=$?" >> $OUT/${TAG}_status.txt
cat $OUT/${TAG}_status.txt
grep -h "SnbError:" $OUT/${TAG}_*.log | sort | uniq -c
