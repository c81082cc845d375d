#!/bin/bash
# per-step kernel, 64-thread CTAs, register cap sweep (warps per SM: 144 -> 14, 128 -> 16, 112 -> 18, 96 -> 20), cold and steady
for r in $1; do
  NAV3D_TPE_BLOCK=64 NAV3D_TPE_REGS=$r python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['steady']; print('regs $r: cold %.1f us (%.3f)  steady %.1f us (%.3f)' % (d['ms_per_step']*1e3, d['roofline']['frac'], s['ms_per_step']*1e3, s['roofline_frac']))"
done
