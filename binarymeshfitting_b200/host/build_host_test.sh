#!/bin/bash
# builds the C++ host-shim test program against the in-tree libbmf_b200.so (rpath = package directory)
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$(dirname "$HERE")"
/usr/bin/g++ -std=c++17 -O2 -Wall -Wno-unused-function -o "$HERE/host_test" "$HERE/host_test.cpp" -L"$PKG" -lbmf_b200 -Wl,-rpath,"$PKG" -lpthread
echo "$HERE/host_test"
