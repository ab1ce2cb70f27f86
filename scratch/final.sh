mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log); tail -3 gpurun_out/pytest_gpu.log
(timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log); cat gpurun_out/smoke.log
(timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?" >> gpurun_out/bench_final.err); python tools/show_bench.py gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err); cut -c1-260 gpurun_out/bench_ref.json; tail -1 gpurun_out/bench_ref.err
