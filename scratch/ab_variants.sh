mkdir -p gpurun_out
export SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_p13.so
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/ab_p13.json 2> gpurun_out/ab_p13.err
echo 'variant [p13]'; python tools/show_bench.py gpurun_out/ab_p13.json | grep -v kernels
export SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_p14.so
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/ab_p14.json 2> gpurun_out/ab_p14.err
echo 'variant [p14]'; python tools/show_bench.py gpurun_out/ab_p14.json | grep -v kernels
