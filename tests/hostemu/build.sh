#!/bin/sh
# TESTS ONLY: compiles the kernel sources with a plain C++ compiler ("warp" of one lane, see
# ima2p_b200/csrc/ima_platform.h) so that the CPU test-suite can exercise the kernel logic against the oracle.
# The package never loads this library; it is not a fallback.
set -e
cd "$(dirname "$0")/../.."
g++ -O2 -std=c++17 -DIMA_HOSTEMU -ffp-contract=off -Wall -Wno-unused-function -Wno-unknown-pragmas -fPIC -shared \
    -x c++ ima2p_b200/csrc/ima_engine.cu -x c++ ima2p_b200/csrc/ima_lmode.cu ima2p_b200/csrc/ima_readu.cpp ima2p_b200/csrc/ima_modelspec.cpp -o tests/hostemu/libima2p_hostemu.so
# the command-line front end against the same test library (CPU tests of the front end's host logic)
g++ -O2 -std=c++17 -Wall ima2p_b200/csrc/ima_frontend.cpp -o tests/hostemu/IMa2p_hostemu -Ltests/hostemu -l:libima2p_hostemu.so -Wl,-rpath,'$ORIGIN'
