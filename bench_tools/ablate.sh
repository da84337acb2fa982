#!/bin/bash
# Per-stage time attribution of the resident kernel: skip one stage at a time (results meaningless, timing only).
for ab in 0 1 2 4 8 16 32 63 62; do
  PRS_RESIDENT_ABLATE=$ab python bench.py --steps 30 --warmup 5 --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ablate', $ab, 'ms_per_step %.4f' % d['ms_per_step'])"
done
