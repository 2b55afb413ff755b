#!/bin/bash
# Tuning aid: tools/probe_frontend.py with the default library and with every variant under tools/variants/.
cd "$(dirname "$0")/.."
cp cpvulkan_b200/csrc/build/libcpvk_cuda.so /tmp/libcpvk_cuda_default.so
echo "default $(python tools/probe_frontend.py 2>&1 | tail -1)"
for f in tools/variants/libcpvk_cuda_*.so; do
  cp "$f" cpvulkan_b200/csrc/build/libcpvk_cuda.so
  echo "$f $(timeout 120 python tools/probe_frontend.py 2>&1 | tail -1)"
done
cp /tmp/libcpvk_cuda_default.so cpvulkan_b200/csrc/build/libcpvk_cuda.so
