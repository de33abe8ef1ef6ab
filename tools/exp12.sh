#!/bin/bash
timeout 2000 python -m pytest tests -q -m gpu 2>&1 | tail -6
CLS="row0_fused self_attn embed update"
echo "== default"; timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== 16 clips"; MSMD_AB_CLIPS=16 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== 1 clip"; MSMD_AB_CLIPS=1 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
touch ubisoft-laforge-msmd_b200/csrc/row0_fused.cu
MSMD_EXTRA_NVCC_FLAGS=-DMSMD_ROW0_TRACE python build.py 2>&1 | tail -1
timeout 300 python tools/ab_step.py row0_fused 2>&1 | grep -E "row0 trace|step" | tail -3
