#!/bin/sh
# TEST INFRASTRUCTURE: the emulation build (build_emu.sh) compiled with AddressSanitizer + UndefinedBehaviorSanitizer, and the
# CPU parity suite run against it.  Global memory, the per-block shared-memory buffers and every staging buffer of the
# library are host heap allocations in this build, so an out-of-range index in a kernel or in the host code is reported
# with a stack.  (compute-sanitizer is not available on the GPU pool; this is the bounds check the kernels get.)
#   tests/emu/run_asan.sh [pytest args]        e.g.  tests/emu/run_asan.sh -k matmul
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
csrc="$root/reference-seal-backend_b200/csrc"
obj="$here/obj_asan"
mkdir -p "$obj"
FLAGS="-fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -g -O1 -pthread -ffp-contract=off -std=c++17 -fPIC -DB200HE_EMU -I$root/tests -I$csrc -Wno-unknown-pragmas"
pids=""
for u in b200he tu_ntt tu_ks tu_moddown; do
    g++ $FLAGS -x c++ -c -o "$obj/$u.o" "$csrc/$u.cu" &
    pids="$pids $!"
done
g++ $FLAGS -c -o "$obj/cuda_shim.o" "$here/cuda_shim.cpp" &
pids="$pids $!"
for p in $pids; do wait "$p"; done
g++ -shared -pthread -fsanitize=address,undefined -o "$obj/libb200he_emu_asan.so" "$obj/b200he.o" "$obj/tu_ntt.o" "$obj/tu_ks.o" "$obj/tu_moddown.o" "$obj/cuda_shim.o"
# the plugin (host mirror of the reference's benchmark classes) with the same sanitizers, on top of that library
back="$root/reference-seal-backend_b200/backend"
make -s -C "$root/reference-seal-backend_b200/hostfhe" >/dev/null 2>&1 || true
g++ -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -g -O1 -std=c++17 -fPIC -fopenmp -I"$back/compat" -I"$back/include" -I"$root/include" \
    -shared -o "$obj/libhebench_seal_backend_emu_asan.so" "$back"/src/engine/*.cpp "$back"/src/benchmarks/*.cpp "$back/compat/hebench/api_bridge/cpp/hebench_cpp.cpp" \
    -L"$obj" -lb200he_emu_asan -L"$root/reference-seal-backend_b200/hostfhe" -lhostfhe -Wl,-rpath,"$obj" -Wl,-rpath,"$root/reference-seal-backend_b200/hostfhe"
# the CPU oracle (the checker itself) with the same sanitizers
g++ -fsanitize=address,undefined -fno-sanitize-recover=undefined -g -O1 -fopenmp -fPIC -std=c++17 -shared -o "$obj/libhe_oracle_asan.so" "$root/oracle/he_oracle.cpp" "$root/oracle/he_oracle_workloads.cpp"
cd "$root"
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
B200HE_EMU_LIB="$obj/libb200he_emu_asan.so" B200HE_EMU_PLUGIN="$obj/libhebench_seal_backend_emu_asan.so" B200HE_ORACLE_LIB="$obj/libhe_oracle_asan.so" python -m pytest tests/test_oracle.py tests/test_emu_parity.py tests/test_workload_parity.py -m "not gpu" -x -q "$@"
