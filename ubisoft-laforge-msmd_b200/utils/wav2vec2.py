"""Drop-in for /root/reference/utils/wav2vec2.py: Wav2Vec2Model with the reference's resampling forward
(wav2vec2.py:71-119; the training-only SpecAugment at :17-53 is out of scope)."""
import torch
import transformers

from .. import _lib


def linear_interpolation(features, input_fps, output_fps, output_len=None):
    """wav2vec2.py:56-62 (host-side helper kept for API parity; the CUDA encoder folds the resampling
    into its loaders)."""
    seq_len = features.shape[2] / float(input_fps)
    if output_len is None:
        output_len = int(seq_len * output_fps)
    return torch.nn.functional.interpolate(features, size=output_len, align_corners=False, mode='linear')


class _AudioEncoderMixin:
    _cfg_cls = None

    @classmethod
    def from_pretrained(cls, name, *a, **k):
        """Released weights when the HF cache has them, random init of the base architecture otherwise
        (no network in the build / bench environment)."""
        try:
            return super().from_pretrained(name, *a, local_files_only=True, **k)
        except Exception:
            return cls(cls._cfg_cls())

    def forward(self, input_values, output_fps=25, attention_mask=None, output_attentions=None,
                output_hidden_states=None, return_dict=None, frame_num=None):
        from ..audio import encode_hidden
        from transformers.modeling_outputs import BaseModelOutput
        if attention_mask is not None:
            raise _lib.MsmdError('msmd_b200 audio encoder: attention_mask is not supported (inference path passes None)')
        hs = encode_hidden(self, input_values, output_fps, frame_num)
        return BaseModelOutput(last_hidden_state=hs, hidden_states=None, attentions=None)

    def extract(self, audio, fps, frame_num, feature_map):
        """MSMD.extract_audio_feature (model.py:250-264) fused: encoder -> 2:1 resample -> Linear(768, d)."""
        from ..audio import extract_audio_feature
        return extract_audio_feature(self, audio, fps, frame_num, feature_map)


class Wav2Vec2Model(_AudioEncoderMixin, transformers.Wav2Vec2Model):
    _cfg_cls = transformers.Wav2Vec2Config
