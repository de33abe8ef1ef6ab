#!/bin/bash
# round-2 experiment batch 1 (one gpurun call): LN rewrite parity + A/B toggles + two-stream overlap + row0 trace
O=gpurun_out
timeout 600 python -m pytest tests/test_denoiser_gpu.py tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -3
CLS="ln1_ln2 ln3 gemm_21312x2048x512 gemm_21312x512x2048_pair gemm_21312x1536x512 gemm_21312x512x512 self_attn cross_attn_row0 row0_fused embed update"
echo "== default (LN rows 6)";           timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== MSMD_LN_ROWS=4";                MSMD_LN_ROWS=4 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== MSMD_GELU_F16X2=1";             MSMD_GELU_F16X2=1 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== MSMD_ROW0_FUSED_MAX_S=192";     MSMD_ROW0_FUSED_MAX_S=192 timeout 300 python tools/ab_step.py $CLS 2>&1 | tail -1
echo "== two streams";                   timeout 400 python tools/two_stream.py 64 1 2 2>&1 | tail -4
echo "== row0 trace at S=192"
touch ubisoft-laforge-msmd_b200/csrc/row0_fused.cu
MSMD_EXTRA_NVCC_FLAGS=-DMSMD_ROW0_TRACE python build.py 2>&1 | tail -1
MSMD_ROW0_FUSED_MAX_S=192 timeout 300 python tools/ab_step.py row0_fused 2>&1 | grep -E "row0 trace|step" | tail -3
