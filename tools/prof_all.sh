#!/bin/bash
# One GPU pass that regenerates the evidence under profiles/ (run under gpurun; outputs land in gpurun_out/).
# Numbers printed under ncu are never bench values: the bench lines come from the plain bench.py runs below.
set -u
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/tests.log 2>&1
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload hires --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_hires.json 2> $O/bench_hires.err
python bench.py --workload batch --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_batch.json 2> $O/bench_batch.err
python tools/time_stages.py 20 > $O/stage_times.txt 2>&1
python tools/time_e2e.py 30 > $O/e2e.txt 2>&1
( timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py ) > $O/memcheck.log 2>&1
# every launch of one step with its device time (cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 16 --csv --log-file $O/launches_step.csv python tools/prof_step.py 2 > /dev/null 2>&1
# full-set capture: single scattering, density order 2, multiple scattering, density order >= 3 (stage-driven pass of prof_step)
ncu --set full --clock-control none --import-source on -k regex:"k_density_main|k_multiple_scattering|k_single_scattering" -s 7 -c 4 -o $O/precompute_full python tools/prof_step.py 2 > /dev/null 2>&1
ncu -i $O/precompute_full.ncu-rep --page raw --csv > $O/precompute_full_raw.csv 2> /dev/null
# speed-of-light + occupancy of every kernel of a step
ncu --section SpeedOfLight --section Occupancy --section LaunchStats --section WarpStateStats --clock-control none -s 16 -c 16 --csv --log-file $O/step_speed_of_light.csv python tools/prof_step.py 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_render_sky -s 4 -c 2 -o $O/render_full python tools/prof_render.py > /dev/null 2>&1
ncu -i $O/render_full.ncu-rep --page raw --csv > $O/render_full_raw.csv 2> /dev/null
python tools/parity_report.py > $O/parity_report.md 2> $O/parity_report.err
tail -3 $O/tests.log; cat $O/bench_default.json; cat $O/bench_hires.json; cat $O/bench_batch.json; cat $O/stage_times.txt
