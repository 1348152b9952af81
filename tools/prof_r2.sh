#!/bin/bash
# Round-2 ncu evidence (run under gpurun; outputs in gpurun_out/, summaries are copied to profiles/ by hand).
# Numbers printed under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
T=${1:-r2}
# every launch of one step with its device time (cold-cache, serialised): compare SHARES with bench.py's ms_per_launch
ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 16 --csv --log-file $O/${T}_launches_step.csv python tools/prof_step.py 2 > /dev/null 2>&1
# full-set capture of the four big precompute kernels (single scattering, density order 2, multiple scattering, density order 3)
ncu --set full --clock-control none --import-source on -k regex:"k_density_main|k_multiple_scattering|k_single_scattering" -s 7 -c 4 -o $O/${T}_precompute_full python tools/prof_step.py 2 > /dev/null 2>&1
ncu -i $O/${T}_precompute_full.ncu-rep --page raw --csv > $O/${T}_precompute_full_raw.csv 2> /dev/null
# the sky evaluation: ground-heavy, space, low-altitude and mixed 4K views
ncu --set full --clock-control none --import-source on -k regex:k_render_sky -s 4 -c 4 -o $O/${T}_render_full python tools/prof_render.py > /dev/null 2>&1
ncu -i $O/${T}_render_full.ncu-rep --page raw --csv > $O/${T}_render_full_raw.csv 2> /dev/null
ls -la $O | grep ${T}_
