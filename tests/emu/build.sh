#!/bin/bash
# Build the TEST-ONLY host emulation of the engine (see backend_emu.cc) into tests/_emu/libmagent_emu.so
set -e
here="$(cd "$(dirname "$0")" && pwd)"
src="$here/../../magent_b200/csrc"
mkdir -p "$here/../_emu"
/usr/bin/g++ -std=c++17 -O2 -g -fPIC -shared -fvisibility=hidden -Wall -Wno-unused-function \
    "$src/engine.cc" "$src/shim.cc" "$src/host_expand.cc" "$here/backend_emu.cc" -pthread -o "$here/../_emu/libmagent_emu.so"
echo "built tests/_emu/libmagent_emu.so"
