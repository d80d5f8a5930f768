#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r3b.log 2>&1
date
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -v "^  \|^$" | cut -c1-300 | tail -12
date
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
date
timeout 600 python bench.py > gpurun_out/r3b_ours.json 2> gpurun_out/r3b_ours.err; tail -2 gpurun_out/r3b_ours.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r3b_ours.json').read().splitlines() if l.startswith('{')][-1]); print('ours', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r3b_ref.json 2> gpurun_out/r3b_ref.err; tail -2 gpurun_out/r3b_ref.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r3b_ref.json').read().splitlines() if l.startswith('{')][-1]); print('ref', d['ms_per_step'], d['value'], d['e2e']['value'])"
date
