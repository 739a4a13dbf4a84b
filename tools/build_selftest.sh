#!/bin/sh
# builds tools/gpu_selftest (torch-free GPU self-test of libar_b200.so); run from the repo root after the library is built
set -e
g++ -std=c++17 -O2 -ffp-contract=off tools/gpu_selftest.cpp -I/usr/local/cuda/include -Iauto_round_b200/csrc \
    -Lauto_round_b200/csrc -lar_b200 -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread \
    -Wl,-rpath,'$ORIGIN/../auto_round_b200/csrc' -o tools/gpu_selftest
echo "built tools/gpu_selftest"
