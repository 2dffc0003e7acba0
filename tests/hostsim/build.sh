#!/bin/bash
# Builds the TEST-ONLY host simulation of the device logic (see hostsim.cpp) into _build/.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
mkdir -p "$here/_build"
CF="-O2 -Wall -Wextra -ffp-contract=off -fPIC -I$root/include"
gcc $CF -c "$root/classpro_b200/host/cpg_model.c" -o "$here/_build/cpg_model.o"
gcc $CF -c "$root/classpro_b200/host/cpg_pack.c"  -o "$here/_build/cpg_pack.o"
g++ $CF -c "$here/hostsim.cpp" -o "$here/_build/hostsim.o"
g++ -shared -o "$here/_build/libhostsim.so" "$here/_build/hostsim.o" "$here/_build/cpg_model.o" "$here/_build/cpg_pack.o" -lm
