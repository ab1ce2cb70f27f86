mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --no-cpu --steps 200 > gpurun_out/bench_h1.json 2>gpurun_out/bench_h1.err
python tools/show_bench.py gpurun_out/bench_h1.json | grep -v kernels; tail -2 gpurun_out/bench_h1.err
