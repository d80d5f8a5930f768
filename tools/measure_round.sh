#!/bin/bash
# The commands behind DESIGN.md section 7 and profiles/*_r2.md (run on one B200 through gpurun; writes gpurun_out/).
# final measurements of the round: launch lists, ncu --set full of one step, bench in all configurations
mkdir -p gpurun_out
exec > gpurun_out/round.log 2>&1
date
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-raster-only --graph off > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2_train.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-raster-only --graph off --loss train > /dev/null 2>&1
date
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_|deform_|tile_|preprocess_|epilogue_|mlp_" -s 57 -c 19 -f -o gpurun_out/prof_r2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-raster-only --graph off 2>&1 | tail -2
date
timeout 600 python bench.py --steps 60 --warmup 8 > gpurun_out/round_ours.json 2> gpurun_out/round_ours.err; tail -2 gpurun_out/round_ours.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 4 > gpurun_out/round_ref.json 2> gpurun_out/round_ref.err; tail -2 gpurun_out/round_ref.err
timeout 600 python bench.py --steps 60 --warmup 8 --loss train --no-cpu-baseline --no-raster-only > gpurun_out/round_ours_train.json 2> gpurun_out/round_ours_train.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 --loss train --no-cpu-baseline --no-raster-only > gpurun_out/round_ref_train.json 2> gpurun_out/round_ref_train.err
timeout 600 python bench.py --steps 60 --warmup 8 --train --no-cpu-baseline --no-raster-only > gpurun_out/round_ours_trainstep.json 2> gpurun_out/round_ours_trainstep.err
for c in C1 C2 C4 C5; do
  timeout 600 python bench.py --config $c --steps 40 --warmup 8 --no-cpu-baseline --no-raster-only > gpurun_out/round_ours_$c.json 2> gpurun_out/round_ours_$c.err
  timeout 600 python bench.py --impl reference --config $c --steps 8 --warmup 3 --no-cpu-baseline --no-raster-only > gpurun_out/round_ref_$c.json 2> gpurun_out/round_ref_$c.err
done
date
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/round_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4f' % d['e2e'].get('ms_per_step', float('nan')) if 'ms_per_step' in d.get('e2e', {}) else '', d.get('raster_only', ''))
    except Exception as e:
        print(f, 'ERR', e)
PY
