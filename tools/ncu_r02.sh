#!/bin/bash
# round-2 ncu evidence (run under gpurun): launch lists + --set full captures of the kernels DESIGN.md quotes
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# (a) every launch of the first 400 of the default bench command, (b) a short sampling run at config-3 shapes
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_bench_launches_first400.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/r02_bench_under_ncu.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/r02_sampler_step_launches.csv python tools/sampler_short.py 64 3 > /dev/null 2>&1
# (c) HBM-bound kernels of the sampling step: cold (ncu default: caches flushed between replays) and warm (--cache-control none:
#     the state they run in inside the replayed step, inputs partly L2-resident)
$NCU --set full -k regex:"ln_kernel|update_kernel|embed_x_kernel|self_attn_tc|cross_attn_row0" -s 30 -c 12 -o $O/r02_step_hbm_cold python tools/sampler_short.py 64 2 > /dev/null 2>&1
$NCU --set full --cache-control none -k regex:"ln_kernel|update_kernel|embed_x_kernel" -s 30 -c 8 -o $O/r02_step_hbm_warm python tools/sampler_short.py 64 2 > /dev/null 2>&1
# (d) fused person-token kernel (16 clips = 48 sequences), rotations
$NCU --set full -k regex:row0_fused -s 8 -c 1 -o $O/r02_row0_fused python tools/sampler_short.py 16 2 > /dev/null 2>&1
$NCU --set full -k regex:rot_kernel -s 9 -c 3 -o $O/r02_rot python tools/rot_one.py > /dev/null 2>&1
ls -la $O/r02_*
