#!/bin/bash
# Profiling aid: device-resident loop of selected stages with differently built libraries (HVO_LIB_PATH).
# usage: tools/variants.sh <stages> <batch> lib1.so lib2.so ...
ST=$1; B=$2; shift 2
P=$(pwd)/a-low-texture-robust-hybrid-feature-based-visual-odometry_b200
for L in "$@"; do
  echo "$L stages $ST batch $B: $(HVO_LIB_PATH=$P/$L python bench.py --device-only --batch $B --steps 5 --warmup 3 --stages $ST | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["value"]), "fps", d["means"])')"
done
