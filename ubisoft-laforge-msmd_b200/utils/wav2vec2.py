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
    def from_pretrained(cls, name, *a, allow_random_init=None, **k):
        """HF ``from_pretrained`` with the caller's arguments (``cache_dir``, ``local_files_only`` ... pass through, as in
        the reference: model.py:94-101).  A checkpoint that cannot be found is an ERROR - a randomly initialised encoder
        produces garbage audio features - unless the caller opts in with ``allow_random_init=True`` or the environment
        variable MSMD_ALLOW_RANDOM_AUDIO_ENCODER=1 (tests / bench, which have no network and no HF cache); the fallback
        then warns loudly.  Only the not-found errors are caught: a corrupt cache or an incompatible checkpoint raises."""
        import os
        import warnings
        if allow_random_init is None:
            allow_random_init = os.environ.get('MSMD_ALLOW_RANDOM_AUDIO_ENCODER', '0') not in ('', '0')
        try:
            return super().from_pretrained(name, *a, **k)
        except (OSError, EnvironmentError) as e:      # HF raises OSError for "not cached / cannot reach the hub"
            if not allow_random_init:
                raise
            warnings.warn(f'msmd_b200: pretrained audio encoder {name!r} is unavailable ({type(e).__name__}); using a RANDOMLY '
                          'INITIALISED base architecture because allow_random_init is set - audio features are '
                          'meaningless unless a full checkpoint is loaded afterwards', RuntimeWarning, stacklevel=2)
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
