#!/bin/bash
# round-2 experiment batch 2: mma.sync person-token kernel for every batch size + warp-per-row update kernel
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_engine_gpu.py tests/test_separate.py tests/test_infer.py -x -q -m gpu 2>&1 | tail -5
CLS="ln1_ln2 ln3 self_attn cross_attn_row0 row0_fused embed update"
echo "== default (fused person-token kernel at every S)";  timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== MSMD_ROW0_FUSED_MAX_S=96 (chain at S=192)";       MSMD_ROW0_FUSED_MAX_S=96 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== 16 clips";  MSMD_AB_CLIPS=16 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== 1 clip";    MSMD_AB_CLIPS=1 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== row0 trace at S=192"
touch ubisoft-laforge-msmd_b200/csrc/row0_fused.cu
MSMD_EXTRA_NVCC_FLAGS=-DMSMD_ROW0_TRACE python build.py 2>&1 | tail -1
timeout 300 python tools/ab_step.py row0_fused 2>&1 | grep -E "row0 trace|step" | tail -3
