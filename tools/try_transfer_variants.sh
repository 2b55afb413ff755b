#!/bin/bash
# Tuning aid: tools/probe_transfer.py (C5 at 8K) with each prebuilt variant of libcpvk_cuda.so under tools/variants/.
cd "$(dirname "$0")/.."
cp cpvulkan_b200/csrc/build/libcpvk_cuda.so /tmp/libcpvk_cuda_default.so
for f in tools/variants/libcpvk_cuda_*.so; do
  cp "$f" cpvulkan_b200/csrc/build/libcpvk_cuda.so
  echo "$f $(python tools/probe_transfer.py 2>/dev/null)"
done
cp /tmp/libcpvk_cuda_default.so cpvulkan_b200/csrc/build/libcpvk_cuda.so
