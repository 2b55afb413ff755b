#!/bin/bash
# Tuning aid: time bench.py's C3/M1 kernels with each prebuilt variant of libcpvk_cuda.so under tools/variants/ (built with
# CPVK_EXTRA_DEFS, see cpvulkan_b200/build.py). Restores the default library afterwards.
cd "$(dirname "$0")/.."
cp cpvulkan_b200/csrc/build/libcpvk_cuda.so /tmp/libcpvk_cuda_default.so
for f in tools/variants/libcpvk_cuda_*.so; do
  cp "$f" cpvulkan_b200/csrc/build/libcpvk_cuda.so
  unset CPVK_RASTER_CTAS
  case "$f" in *ctas5*) export CPVK_RASTER_CTAS=5;; *ctas4*) export CPVK_RASTER_CTAS=4;; *ctas3*) export CPVK_RASTER_CTAS=3;; *ctas2*) export CPVK_RASTER_CTAS=2;; esac
  for i in 1 2; do
    python bench.py --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['value'],1), d['kernel_ms_rank0'])"
  done
  # the C4-shaped draw (100 of its quads) as well when asked: CPVK_VARIANTS_C4=1
  if [ -n "${CPVK_VARIANTS_C4:-}" ]; then
    CPVK_BENCH_C4_QUADS=100 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', 'C4', round(d['value'],2), 'Gfragments/s')"
  fi
done
cp /tmp/libcpvk_cuda_default.so cpvulkan_b200/csrc/build/libcpvk_cuda.so
