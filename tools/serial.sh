#!/bin/bash
run() { echo "$1: $(env $2 python bench.py --device-only --frames $3 --batch $4 --lanes $5 --steps 3 --warmup 2 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],2), "ms", round(d["value"]), "fps")' 2>&1)"; }
run "serial1 4736 lanes1"        "HVO_FRAME_SERIAL=1" 9472 4736 1
run "serial2 4736 lanes1"        "HVO_FRAME_SERIAL=2" 9472 4736 1
run "serial3 4736 lanes1"        "HVO_FRAME_SERIAL=3" 9472 4736 1
run "serial2 3552 lanes1"        "HVO_FRAME_SERIAL=2" 10656 3552 1
run "serial3 3552 lanes1"        "HVO_FRAME_SERIAL=3" 10656 3552 1
run "serial2 2368 lanes1"        "HVO_FRAME_SERIAL=2" 9472 2368 1
