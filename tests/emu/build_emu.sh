#!/bin/sh
# TEST INFRASTRUCTURE: compiles the product's CUDA sources as host C++ (fibers emulate CUDA threads)
# so kernel index arithmetic can be checked against the oracle without a GPU.  Not shipped, not a fallback.
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
csrc="$root/reference-seal-backend_b200/csrc"
obj="$here/obj"
mkdir -p "$obj"
FLAGS="$EMU_EXTRA -O2 -pthread -ffp-contract=off -std=c++17 -fPIC -DB200HE_EMU -I$root/tests -I$csrc -Wno-unknown-pragmas"
pids=""
for u in b200he tu_ntt tu_ks tu_moddown; do
    g++ $FLAGS -x c++ -c -o "$obj/$u.o" "$csrc/$u.cu" &
    pids="$pids $!"
done
g++ $FLAGS -c -o "$obj/cuda_shim.o" "$here/cuda_shim.cpp" &
pids="$pids $!"
for p in $pids; do wait "$p"; done
g++ -shared -pthread -o "$here/libb200he_emu.so" "$obj/b200he.o" "$obj/tu_ntt.o" "$obj/tu_ks.o" "$obj/tu_moddown.o" "$obj/cuda_shim.o"
