"""Drop-in for /root/reference/utils/hubert.py: HubertModel with the reference's resampling forward.

The class is a PARAMETER HOLDER with HF ``transformers.HubertModel``'s state_dict layout (so released
checkpoints load unchanged); the arithmetic (conv front-end, feature projection, positional conv,
12 post-LN encoder layers, hubert.py:13-51) runs in the CUDA audio encoder (csrc/audio_*.cu).
"""
import transformers

from .wav2vec2 import _AudioEncoderMixin


class HubertModel(_AudioEncoderMixin, transformers.HubertModel):
    _cfg_cls = transformers.HubertConfig
