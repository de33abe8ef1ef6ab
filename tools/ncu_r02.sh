#!/bin/bash
# round-2 final evidence (one gpurun call): bench lines of every single-GPU workload, ncu launch lists, --set full captures
# of the kernels DESIGN.md quotes.  Numbers printed under ncu are never bench values.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# ---- bench lines (not under a profiler)
python bench.py > $O/r02f_bench_default.json 2> $O/r02f_bench_default.err
python bench.py --workload flame > $O/r02f_bench_flame.json 2>/dev/null
python bench.py --workload latency1 --no-cpu-baseline > $O/r02f_bench_latency1.json 2>/dev/null
python bench.py --workload wav2vec2_60s --no-cpu-baseline --steps 1 > $O/r02f_bench_w2v.json 2>/dev/null
python bench.py --impl reference --steps 1 --warmup 1 > $O/r02f_bench_reference.json 2>/dev/null
# ---- launch lists: the first 400 launches of the default bench command; one sampling step at configuration-3 shapes
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02f_bench_launches_first400.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/r02f_sampler_step_launches.csv python tools/sampler_short.py 64 3 > /dev/null 2>&1
$NCU --cache-control none --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/r02f_sampler_step_launches_warm.csv python tools/sampler_short.py 64 3 > /dev/null 2>&1
# ---- --set full: GEMMs (clusters of 4), FLAME (two column tiles per A block), the per-step small kernels in the state they run in
$NCU --set full -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/r02f_gemm_ff1 python tools/gemm_one.py ff1 > /dev/null 2>&1
$NCU --set full -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/r02f_gemm_qkv python tools/gemm_one.py qkv > /dev/null 2>&1
$NCU --set full -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/r02f_gemm_ff2 python tools/gemm_one.py ff2 > /dev/null 2>&1
$NCU --set full -k regex:flame_tc_kernel -s 3 -c 1 -o $O/r02f_flame python tools/flame_time.py > /dev/null 2>&1
$NCU --set full --cache-control none -k regex:"ln_kernel|update_kernel|embed_x|row0_fused|self_attn_tc" -s 40 -c 10 -o $O/r02f_step_small_warm python tools/sampler_short.py 64 2 > /dev/null 2>&1
ls -la $O/r02f_*
