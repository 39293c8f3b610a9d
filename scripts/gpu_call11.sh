#!/bin/bash
# round 2, call 11: ncu --set full on one launch of each of the heaviest GEMM shapes (what bounds them?)
mkdir -p gpurun_out
for i in 0 1 2 3 4 7 10 5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm --launch-skip 12 --launch-count 1 \
    -o gpurun_out/r02_gemm_shape$i -f python scripts/bench_gemm.py $i > gpurun_out/ncu_gemm_$i.log 2>&1
  echo "shape $i rc=$?"; grep "TF/s" gpurun_out/ncu_gemm_$i.log
done
python scripts/bench_gemm.py 0 1 2 3 4 5 6 7 10 11 2>&1 | grep "TF/s"
ls -la gpurun_out/*.ncu-rep
