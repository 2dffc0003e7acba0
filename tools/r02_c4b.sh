#!/bin/bash
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
rm -rf /tmp/big_c4; python tools/run_big.py c4 --genome-mb 50 --keep /tmp/big_c4 > $O/big_c4_$TAG.json 2> $O/big_c4_$TAG.err; tail -1 $O/big_c4_$TAG.json | cut -c1-300
python tools/trace_file_flips.py /tmp/big_c4 60 $O/flipreads_$TAG.npz > $O/flips_c4_$TAG.json 2> $O/flips_c4_$TAG.err; grep -c "host_build_of_device_code_equals_reference\": true" $O/flips_c4_$TAG.json; grep -c "host_build_of_device_code_equals_reference\": false" $O/flips_c4_$TAG.json; rm -rf /tmp/big_c4
