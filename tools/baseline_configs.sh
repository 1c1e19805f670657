#!/bin/sh
# BASELINE.json configs through the plugin (libhebench_seal_backend.so driven by the mini harness in HEBench order),
# at their stated shapes; writes one CSV per config under gpurun_out/.  Usage: tools/baseline_configs.sh [tag]
tag=${1:-r1}
H=./reference-seal-backend_b200/backend/mini_harness
P=reference-seal-backend_b200/backend/libhebench_seal_backend.so
run() { name=$1; shift; echo "== $name: $*"; $H --backend_lib_path $P --random_seed 1234 --csv gpurun_out/configs_${tag}_$name.csv "$@" 2>&1 | grep -E "operate|Failed|Total|Error|rror" ; }
run C1 --filter "EltwiseMultiply BFV Offline" --n 100 --samples 10,10
run C2add --filter "EltwiseAdd CKKS Offline" --n 1000 --samples 100,10
run C2mul --filter "EltwiseMultiply CKKS Offline" --n 1000 --samples 100,10
run C3 --filter "DotProduct CKKS Offline" --n 100 --poly 16384 --samples 100,100
run C4val --filter "MatrixMultiply CKKS Latency other=0" --dims 100,100,100 --poly 16384 --depth 6 --iterations 1
run C4cba --filter "MatrixMultiply CKKS Latency other=1" --dims 20,20,20 --poly 16384 --depth 6 --iterations 1
run C5 --filter "LogisticRegression_PolyD3 CKKS Offline" --poly 32768 --batch 1024 --iterations 1
