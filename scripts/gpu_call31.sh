#!/bin/bash
set -u
O=gpurun_out/c31; mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --e2e-steps 2 --no-cpu --no-parity --no-slab --no-ensemble --profile-steps 2"
for v in base nod2h chunk8 chunk128; do
  case $v in base) E="";; nod2h) E="VCH_DEBUG_NO_D2H=1";; chunk8) E="VCH_STREAM_CHUNK=8";; chunk128) E="VCH_STREAM_CHUNK=128";; esac
  env VCH_DEBUG=1 $E timeout 600 $B > $O/$v.json 2> $O/$v.err
  echo "== $v"; grep "streamed host path" $O/$v.err | tail -1
  python -c "
import json; d=json.loads(open('$O/$v.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
  rm -f $O/$v.err
done
