#!/usr/bin/env python
"""GPU probe: distribution of ||R|| / floor_est at the Newton stop over a 1024^2 x M trajectory (uncontrolled + controlled).
Run with VCH_DEBUG=1 and parse the library's stderr lines.  usage: VCH_DEBUG=1 python scripts/newton_floor_probe.py [N] [M] 2> log"""
import os, sys
sys.argv = [sys.argv[0]] + sys.argv[1:]
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_vs_strict as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
a = P.run(N, M, strict=False)
print(a["fwd_stats"], a["it_stats"])
