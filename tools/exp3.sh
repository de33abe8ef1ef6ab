#!/bin/bash
# round-2 experiment batch 3: L2 prefetch of the cross-attention caches under self-attention; person-token kernel phase 4 without statistics exchange
O=gpurun_out
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_engine_gpu.py tests/test_style.py -x -q -m gpu 2>&1 | tail -3
CLS="ln1_ln2 ln3 self_attn row0_fused embed update"
echo "== default";               timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== MSMD_L2_PREFETCH=0";    MSMD_L2_PREFETCH=0 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== chain, prefetch on";    MSMD_ROW0_FUSED_MAX_S=96 timeout 300 python tools/ab_step.py $CLS cross_attn_row0 2>&1 | tail -1
echo "== 16 clips";  MSMD_AB_CLIPS=16 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== 1 clip";    MSMD_AB_CLIPS=1 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
ncu --clock-control none --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/r02b_sampler_step_launches.csv python tools/sampler_short.py 64 3 > /dev/null 2>&1
python tools/launch_agg.py $O/r02b_sampler_step_launches.csv
echo "== row0 trace at S=192"
touch ubisoft-laforge-msmd_b200/csrc/row0_fused.cu
MSMD_EXTRA_NVCC_FLAGS=-DMSMD_ROW0_TRACE python build.py 2>&1 | tail -1
timeout 300 python tools/ab_step.py row0_fused 2>&1 | grep -E "row0 trace|step" | tail -3
