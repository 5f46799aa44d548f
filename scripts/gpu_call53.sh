#!/bin/bash
set -u
O=gpurun_out/c53; mkdir -p $O
run() { name=$1; shift; env "$@" PROBE_REPS=16 timeout 200 python scripts/determinism_probe.py 1024 6 3 > $O/$name.txt 2>&1; echo "== $name ($*)"; grep -E "jacobian_solve|newton step|vs run|residual:" $O/$name.txt; }
run A_notma VCH_COLS_TMA=0
run B_notma_nographs VCH_COLS_TMA=0 VCH_NO_GRAPHS=1
run C_notma_nofused VCH_COLS_TMA=0 VCH_FUSED_SOLVE=0
run D_notma_notiled VCH_COLS_TMA=0 VCH_TILED=0
run E_notma_bicg7 VCH_COLS_TMA=0 VCH_BICG6=0
run F_notma_nohalf VCH_COLS_TMA=0 VCH_NO_HALF_EXIT=1
run G_default VCH_DEBUG=0
run H_radix8 VCH_FFT16=0
