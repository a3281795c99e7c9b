#!/bin/bash
# usage: tools/bench_variant.sh lib1.so lib2.so ...   (GPU box) - headline bench line for alternative builds
for lib in "$@"; do
  RVLP_LIB=$PWD/build_variants/$lib python bench.py --no-extras --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', 'value %.1f G/s  %.3f ms  e2e %.1f G/s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9))"
done
