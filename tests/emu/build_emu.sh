#!/bin/sh
# TEST INFRASTRUCTURE: builds the product's .cu sources as plain C++ against the SIMT emulator.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
g++ -O1 -g -std=c++17 -fPIC -rdynamic -shared -pthread -DLB_SIMT_EMU -I"$HERE" -Wall -Wno-unused-function -Wno-unknown-pragmas \
    -x c++ "$ROOT/loro_b200/csrc/engine.cu" -x c++ "$HERE/simt_emu.cpp" -o "$HERE/libloro_b200_emu.so"
