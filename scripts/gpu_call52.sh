#!/bin/bash
set -u
O=gpurun_out/c52; mkdir -p $O
timeout 300 ./scripts/cols_determinism 16 > $O/cols_determinism.txt 2>&1; echo rc=$? >> $O/cols_determinism.txt
cat $O/cols_determinism.txt
