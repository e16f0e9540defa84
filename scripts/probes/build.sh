#!/bin/bash
# builds the standalone probes in-tree (the binary travels to the GPU box with gpurun; it is git-ignored)
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -o tc_probe.bin tc_probe.cu
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -o mma_probe.bin mma_probe.cu
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -o hbm_probe.bin hbm_probe.cu
