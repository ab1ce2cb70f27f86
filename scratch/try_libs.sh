#!/bin/bash
# run the bench with alternative builds of the library (tuning experiments)
for v in "$@"; do
  cp openmm-lab_b200/csrc/libslicednb_b200.so /tmp/keep.so
  cp scratch/libs/lib$v.so openmm-lab_b200/csrc/libslicednb_b200.so
  python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/b_lib$v.log 2>&1
  cp /tmp/keep.so openmm-lab_b200/csrc/libslicednb_b200.so
done
