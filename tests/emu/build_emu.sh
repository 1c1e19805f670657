#!/bin/sh
# TEST INFRASTRUCTURE: compiles the product's CUDA sources as host C++ (fibers emulate CUDA threads)
# so kernel index arithmetic can be checked against the oracle without a GPU.  Not shipped, not a fallback.
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
g++ -O2 -pthread -ffp-contract=off -std=c++17 -fPIC -shared -DB200HE_EMU -x c++ -I"$root/tests" -I"$root/reference-seal-backend_b200/csrc" \
    -Wno-unknown-pragmas -o "$here/libb200he_emu.so" \
    "$root/reference-seal-backend_b200/csrc/b200he.cu" "$here/cuda_shim.cpp"
