#!/bin/bash
# What the host side gives N GPUs together: frame-sized copy-engine transfers on 1, 2, 4, 8 devices at once (run under gpurun --gpus 8).
cd "$(dirname "$0")/.."
[ -x scripts/bin/pcie_probe ] || { mkdir -p scripts/bin && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart static -o scripts/bin/pcie_probe scripts/pcie_probe.cu; }
for mode in both h2d d2h; do
  for n in 1 2 4 8; do
    echo "== $mode on $n GPUs at once"
    for ((i = 0; i < n; i++)); do CUDA_VISIBLE_DEVICES=$i timeout 60 scripts/bin/pcie_probe sustain $mode 1.5 & done
    wait
  done
done
