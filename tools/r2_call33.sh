#!/usr/bin/env bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -k "$1" > gpurun_out/t_$name.log 2>&1; echo "exit $?"; tail -1 gpurun_out/t_$name.log; }
IEF_TC_VERSION=1 run tc_v1 "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_CROSS_TC=0 run cross_mma_only "cross_attention"
IEF_CROSS_TC_EDIT=0 run cross_edit_on_mma "cross_attention"
run all ""
