"""Host side of the CUDA audio encoder (csrc/audio*.cu): utils/hubert.py:13-51, utils/wav2vec2.py:71-119,
model.py:250-264.  One engine per encoder module, sized on demand; no HF / PyTorch compute fallback."""
import ctypes as C

import torch

from . import _lib


class _AudioEngine:
    def __init__(self, max_clips, max_samples, d_out, device):
        h = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().msmd_audio_create(max_clips, max_samples, d_out, idx, C.byref(h)))
        self._h, self.cap, self.d_out, self.device, self.key = h, (max_clips, max_samples), d_out, device, None

    def __del__(self):
        try:
            _lib.lib().msmd_audio_destroy(self._h)
        except Exception:
            pass


def _engine(module, N, n_samples, feature_map, device):
    if device.type != 'cuda':
        raise _lib.MsmdError('msmd_b200 audio encoder needs CUDA tensors (no CPU path)')
    d_out = feature_map.out_features if feature_map is not None else 512
    eng = module.__dict__.get('_msmd_engine')
    if eng is None or eng.device != device or eng.cap[0] < N or eng.cap[1] < n_samples or eng.d_out != d_out:
        cap = (max(N, eng.cap[0] if eng else 0), max(n_samples, eng.cap[1] if eng else 0))
        eng = _AudioEngine(cap[0], cap[1], d_out, device)
        module.__dict__['_msmd_engine'] = eng
    sd = {'audio_encoder.' + k: v for k, v in module.state_dict().items() if v.is_floating_point()}
    if feature_map is not None:
        sd['audio_feature_map.weight'], sd['audio_feature_map.bias'] = feature_map.weight, feature_map.bias
    key = tuple((k, v.data_ptr(), v._version) for k, v in sd.items())
    if key != eng.key:
        items = [(k, v.detach().float().contiguous()) for k, v in sd.items()]
        n = len(items)
        names = (C.c_char_p * n)(*[k.encode() for k, _ in items])
        ptrs = (C.c_void_p * n)(*[v.data_ptr() for _, v in items])
        numel = (C.c_int64 * n)(*[v.numel() for _, v in items])
        with torch.cuda.device(device):
            torch.cuda.synchronize()
            _lib.check(_lib.lib().msmd_audio_load_weights(eng._h, names, ptrs, numel, n))
        eng.key = key
    return eng


@torch.no_grad()
def encode_hidden(module, input_values, output_fps, frame_num):
    """hubert.py:13-51 with ALREADY padded input (the wrapper receives pad_audio(audio), model.py:257).
    The library folds pad_audio itself, so this entry point un-pads nothing: it is given raw clips by
    extract_audio_feature; called directly it treats input_values as raw audio without padding rules."""
    raise _lib.MsmdError('msmd_b200: call MSMD.extract_audio_feature (model.py:250) - the standalone wrapper '
                         'forward expects pre-padded audio, which the fused CUDA loader does not take')


@torch.no_grad()
def extract_audio_feature(module, audio, fps, frame_num, feature_map, return_hidden=False):
    """model.py:250-264: audio [N, n] -> [N, frame_num, d] (encoder runs at 2*frame_num frames)."""
    x = _lib.as_f32c(audio)
    N, n = x.shape
    eng = _engine(module, N, n, feature_map, x.device)
    feat = torch.empty((N, frame_num, eng.d_out), device=x.device)
    hidden = torch.empty((N, 2 * frame_num, 768), device=x.device) if return_hidden else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().msmd_audio_encode(eng._h, _lib.dev_ptr(x), N, n, int(fps), 2 * frame_num,
                                                _lib.dev_ptr(hidden), frame_num, _lib.dev_ptr(feat),
                                                _lib.stream_ptr()))
    return (feat, hidden) if return_hidden else feat
