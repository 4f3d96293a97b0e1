#!/bin/bash
# each case in its own process: a trap in one kernel must not poison the others
cd "$(dirname "$0")/.."
for c in gemm32 gemm64 gemm128 gemm256 gemm512 k1_cin32 k3_s1 k3_s2_cin32 k1_s2 k3_res_relu k3_leaky concat_up_k3 concat_up_k1 concat_up_768 deepk_512 many_tiles; do
  timeout 120 python tools/conv_debug.py $c 2>&1 | grep -E "CASE|bad|m=|Error|error" | head -12
  echo "  rc=$?"
done
