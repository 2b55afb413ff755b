#!/bin/bash
# Tuning aid: time bench.py's C3/M1 kernels with each prebuilt variant of libcpvk_cuda.so under tools/variants/ (built with
# CPVK_EXTRA_DEFS, see cpvulkan_b200/build.py). Restores the default library afterwards.
cd "$(dirname "$0")/.."
cp cpvulkan_b200/csrc/build/libcpvk_cuda.so /tmp/libcpvk_cuda_default.so
for f in tools/variants/libcpvk_cuda_*.so; do
  cp "$f" cpvulkan_b200/csrc/build/libcpvk_cuda.so
  for i in 1 2; do
    python bench.py --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['value'],1), d['kernel_ms_rank0']['raster'])"
  done
done
cp /tmp/libcpvk_cuda_default.so cpvulkan_b200/csrc/build/libcpvk_cuda.so
