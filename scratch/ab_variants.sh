mkdir -p gpurun_out
export SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_sb16.so
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/ab_sb16.json 2> gpurun_out/ab_sb16.err
echo 'variant [sb16]'; python tools/show_bench.py gpurun_out/ab_sb16.json | grep -v kernels
export SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_sb8.so
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/ab_sb8.json 2> gpurun_out/ab_sb8.err
echo 'variant [sb8]'; python tools/show_bench.py gpurun_out/ab_sb8.json | grep -v kernels
export SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200_sb1616.so
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/ab_sb1616.json 2> gpurun_out/ab_sb1616.err
echo 'variant [sb1616]'; python tools/show_bench.py gpurun_out/ab_sb1616.json | grep -v kernels
