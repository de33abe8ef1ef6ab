"""Host side of the CUDA audio encoder (utils/hubert.py:13-51, utils/wav2vec2.py:71-119,
model.py:250-264).  Not built yet in this revision: calls fail loudly (no CPU / HF fallback)."""
from . import _lib


def encode_hidden(module, input_values, output_fps, frame_num):
    raise _lib.MsmdError('msmd_b200: the CUDA audio encoder is not built yet')


def extract_audio_feature(module, audio, fps, frame_num, feature_map):
    raise _lib.MsmdError('msmd_b200: the CUDA audio encoder is not built yet')
