#!/bin/bash
# Profiling aid: does the front-end's step time equal the longest pipeline (streams overlap) or the sum?
run() { python bench.py --device-only --batch $1 --steps 3 --warmup 3 --stages $2 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],2), "ms", round(d["value"]), "fps")'; }
for B in "$@"; do
  for st in 15 6 7 14; do
    echo "B=$B prio on  stages $st: $(run $B $st)"
    echo "B=$B prio off stages $st: $(HVO_FRAME_PRIORITIES=0 run $B $st)"
  done
done
