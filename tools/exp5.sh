#!/bin/bash
# round-2 experiment batch 5: tensor-core embedding, update kernel with staged parameters, LayerNorm default = one row per warp
O=gpurun_out
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_engine_gpu.py tests/test_separate.py tests/test_infer.py -x -q -m gpu 2>&1 | tail -3
CLS="ln1_ln2 ln3 embed update row0_fused"
echo "== default";               timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== MSMD_EMBED_MMA=0";      MSMD_EMBED_MMA=0 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== MSMD_LN_ROWS=1";        MSMD_LN_ROWS=1 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== default again";         timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== 1 clip";                MSMD_AB_CLIPS=1 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== 1 clip MSMD_LN_ROWS=1"; MSMD_LN_ROWS=1 MSMD_AB_CLIPS=1 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/r02d_sampler_step_launches_warm.csv python tools/sampler_short.py 64 3 > /dev/null 2>&1
python tools/launch_agg.py $O/r02d_sampler_step_launches_warm.csv
