mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
SNB_NO_PACKED=1 timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_scalar.log 2>&1; tail -3 gpurun_out/pytest_gpu_scalar.log
for t in packed scalar; do
  if [ $t = scalar ]; then export SNB_NO_PACKED=1; else unset SNB_NO_PACKED; fi
  timeout 200 python bench.py --no-cpu --steps 100 > gpurun_out/bench_$t.json 2>gpurun_out/bench_$t.err
  echo "variant [$t]"; python tools/show_bench.py gpurun_out/bench_$t.json | grep -v kernels
done
