#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/final2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/final2_tests.log
tail -4 gpurun_out/final2_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
