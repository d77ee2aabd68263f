#!/bin/bash
# Build build/libard_trace.so: the product library with the per-role clock64() stamps compiled into the fused FFN kernels
# (-DARD_FFN_TRACE), for tools/ffn_trace.py and tools/ffw_trace.py. Needs the product objects (python __graft_entry__.py) first.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/trace
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DARD_FFN_TRACE -DARD_AB_TRACE"
nvcc $F -c audio_residual_b200/csrc/ffn_fused.cu -o build/trace/ffn_fused.o &
nvcc $F -c audio_residual_b200/csrc/ffn_wide.cu -o build/trace/ffn_wide.o &
nvcc $F -c audio_residual_b200/csrc/attn_block.cu -o build/trace/attn_block.o &
wait
nvcc --shared -gencode arch=compute_100a,code=sm_100a -o build/libard_trace.so build/trace/ffn_fused.o build/trace/ffn_wide.o build/trace/attn_block.o \
    $(ls build/*.cu.o | grep -v "ffn_fused\|ffn_wide\|attn_block")
ls -la build/libard_trace.so
