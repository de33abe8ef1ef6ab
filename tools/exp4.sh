#!/bin/bash
# round-2 experiment batch 4: LayerNorm kernel A/B on one box (round-1 kernel vs the 6-rows-per-warp one), single-wave update kernel
O=gpurun_out
timeout 600 python -m pytest tests/test_denoiser_gpu.py tests/test_separate.py -x -q -m gpu 2>&1 | tail -2
CLS="ln1_ln2 ln3 update"
for rep in 1 2; do
echo "== LN new (6 rows/warp)";   timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== LN round-1 kernel";      MSMD_LN_ROWS=0 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
done
echo "== LN 4 rows/warp";         MSMD_LN_ROWS=4 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/r02c_sampler_step_launches_warm.csv python tools/sampler_short.py 64 3 > /dev/null 2>&1
python tools/launch_agg.py $O/r02c_sampler_step_launches_warm.csv
