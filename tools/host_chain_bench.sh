#!/bin/bash
# builds and runs tools/host_chain_bench.cpp (no GPU needed): host cost per op of the slots runtime
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_build
gcc -O2 -Iinclude -c tests/support/hostmem_backend.c -o tools/_build/hostmem_backend.o
g++ -std=c++17 -O2 -Iinclude -o tools/_build/host_chain_bench tools/host_chain_bench.cpp tests/native/dataplane_stubs.cpp \
    lzy_b200/csrc/host/*.cpp tools/_build/hostmem_backend.o -lpthread
# 5th argument "socket": the matcher calls of every slot cross the box-local transport (a unix socket under /tmp)
SOCK=""
if [ "${5:-}" = "socket" ]; then SOCK="/tmp/lzyb-hcb-$$.sock"; fi
for i in 1 2 3; do tools/_build/host_chain_bench "${1:-1000}" "${2:-8}" "${3:-1}" "${4:-0}" $SOCK; done
