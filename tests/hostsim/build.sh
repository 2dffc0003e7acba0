#!/bin/bash
# Builds the TEST-ONLY host simulations of the device logic (see hostsim.cpp) into _build/:
# libhostsim.so (warp width 1) and libhostsim32.so (32 host threads per warp).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
mkdir -p "$here/_build"
CF="-O2 -Wall -Wextra -ffp-contract=off -fPIC -I$root/include"
gcc $CF -c "$root/classpro_b200/host/cpg_model.c" -o "$here/_build/cpg_model.o"
gcc $CF -c "$root/classpro_b200/host/cpg_pack.c"  -o "$here/_build/cpg_pack.o"
g++ $CF -DCPG_HOSTSIM=1 -c "$here/hostsim.cpp" -o "$here/_build/hostsim.o"
g++ -shared -o "$here/_build/libhostsim.so" "$here/_build/hostsim.o" "$here/_build/cpg_model.o" "$here/_build/cpg_pack.o" -lm
g++ $CF -DCPG_HOSTSIM=32 -c "$here/hostsim.cpp" -o "$here/_build/hostsim32.o"
g++ -shared -o "$here/_build/libhostsim32.so" "$here/_build/hostsim32.o" "$here/_build/cpg_model.o" "$here/_build/cpg_pack.o" -lm -lpthread
# the command-line programs on top of a TEST-ONLY stand-in for the device (fakedev.cpp): the host
# side of the product (reader, batching, packing pool, workers, ordered writer) for the CPU suite
g++ $CF -c "$here/fakedev.cpp" -o "$here/_build/fakedev.o"
gcc $CF -I"$root/classpro_b200/host" -c "$root/classpro_b200/host/classpro_main.c" -o "$here/_build/classpro_main.o"
gcc $CF -I"$root/classpro_b200/host" -c "$root/classpro_b200/host/cpg_prof2class.c" -o "$here/_build/cpg_prof2class.o"
g++ -o "$here/_build/ClassPro" "$here/_build/classpro_main.o" "$here/_build/fakedev.o" "$here/_build/cpg_model.o" "$here/_build/cpg_pack.o" -lz -lpthread -lm
g++ -o "$here/_build/prof2class" "$here/_build/cpg_prof2class.o" "$here/_build/fakedev.o" "$here/_build/cpg_model.o" "$here/_build/cpg_pack.o" -lz -lpthread -lm
# TEST-ONLY host build of the profile producer's element functions (countsim.cpp)
g++ $CF -shared -o "$here/_build/libcountsim.so" "$here/countsim.cpp"
g++ $CF -DCOUNTSIM_AS_ABI -c "$here/countsim.cpp" -o "$here/_build/countsim_abi.o"
gcc $CF -I"$root/classpro_b200/host" -c "$root/classpro_b200/host/cpg_profiler.c" -o "$here/_build/cpg_profiler.o"
g++ -o "$here/_build/profiler" "$here/_build/cpg_profiler.o" "$here/_build/countsim_abi.o" "$here/_build/cpg_pack.o" -lz
# TEST-ONLY: the key-range-pass kernels of the profile producer on host threads (passemu.cpp)
g++ $CF -shared -o "$here/_build/libpassemu.so" "$here/passemu.cpp" -lpthread
