"""Drop-in for /root/reference/style_encoder.py: get_style_encoder / StyleEncoder_VAE2.

The module holds the parameters under the reference's state_dict keys; forward / sample
run in csrc/style.cu behind the C ABI (msmd_style_*).  The Gaussian noise stays torch's
(``torch.randn_like``, drawn where the reference draws it: once in forward, twice in sample -
style_encoder.py:200, :212), so seeding behaves the same.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .utils.model_common import PositionalEncoding


def get_style_encoder(args, style_encoder_model_style="diffposetalk"):
    """style_encoder.py:7-12 (only "vae2" returns a model there too)."""
    if style_encoder_model_style == "vae2":
        return StyleEncoder_VAE2(args)
    return None


def _slots(n, mods):
    """nn.Sequential of n slots: the given {index: module} entries, Identity elsewhere (keeps the reference's
    numbering of input_layers / output_layers without its Permute / Dropout / ELU helper modules)."""
    return nn.Sequential(*[mods.get(i, nn.Identity()) for i in range(n)])


class StyleEncoder_VAE2(nn.Module):
    def __init__(self, args) -> None:
        super().__init__()
        self.input_dim = 67
        if args.dataset_type[:9] == 'HDTF_TFHP' or args.dataset_type == "flame_mead_ravdess":
            self.input_dim = 54
        self.motion_coef_dim = self.input_dim
        self.conv_feature_dim = 512
        self.output_size = args.d_style * 2
        d, o = self.conv_feature_dim, self.output_size
        self.input_layers = _slots(12, {1: nn.Conv1d(self.input_dim, d, 3, padding=1), 5: nn.LayerNorm(d),
                                        7: nn.Conv1d(d, d, 3, padding=1), 11: nn.LayerNorm(d)})
        self.PE = PositionalEncoding(d)
        self.encoder = nn.TransformerEncoderLayer(d_model=d, nhead=8, dim_feedforward=d, activation='gelu',
                                                  batch_first=True)                      # parameter holder only
        self.output_layers = _slots(9, {1: nn.Conv1d(d, o, 3, padding=1), 5: nn.LayerNorm(o),
                                        7: nn.Conv1d(o, o, 3, padding=1)})
        self._h = None
        self._key = None
        self._cap = (0, 0)

    def _handle(self, N, L, device):
        if device.type != 'cuda':
            raise _lib.MsmdError('msmd_b200 style encoder needs CUDA tensors (no CPU path)')
        lib = _lib.lib()
        if self._h is None or self._cap[0] < N or self._cap[1] < L or self._dev != device:
            if self._h is not None:
                lib.msmd_style_destroy(self._h)
            h = C.c_void_p()
            cap = (max(N, self._cap[0]), max(L, self._cap[1]))
            idx = device.index if device.index is not None else torch.cuda.current_device()
            _lib.check(lib.msmd_style_create(self.input_dim, self.conv_feature_dim, self.output_size // 2, cap[0], cap[1],
                                             idx, C.byref(h)))
            self._h, self._cap, self._dev, self._key = h, cap, device, None
        sd = {k: v for k, v in self.state_dict().items() if v.is_floating_point()}
        key = tuple((k, v.data_ptr(), v._version) for k, v in sd.items())
        if key != self._key:
            items = [(k, v.detach().float().contiguous()) for k, v in sd.items()]
            n = len(items)
            names = (C.c_char_p * n)(*[k.encode() for k, _ in items])
            ptrs = (C.c_void_p * n)(*[v.data_ptr() for _, v in items])
            numel = (C.c_int64 * n)(*[v.numel() for _, v in items])
            with torch.cuda.device(device):
                torch.cuda.synchronize()
                _lib.check(lib.msmd_style_load_weights(self._h, names, ptrs, numel, n))
            self._key = key
        return self._h

    @torch.no_grad()
    def _stats(self, motion_coef):
        N, L, _ = motion_coef.shape
        x = _lib.as_f32c(motion_coef)
        h = self._handle(N, L, x.device)
        mu = torch.empty((N, self.output_size // 2), device=x.device)
        logvar = torch.empty_like(mu)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().msmd_style_encode(h, _lib.dev_ptr(x), N, L, None, None, _lib.dev_ptr(mu),
                                                    _lib.dev_ptr(logvar), _lib.stream_ptr()))
        return mu, logvar

    def forward(self, motion_coef, do_sample=False):
        """style_encoder.py:178-208: returns mu + eps*std if do_sample else (mu + eps*std, mu, logvar)."""
        mu, logvar = self._stats(motion_coef)
        std = torch.exp(0.5 * logvar)
        out = mu + torch.randn_like(std) * std
        return out if do_sample else (out, mu, logvar)

    def sample(self, motion_coef):
        """style_encoder.py:209-213: draws eps twice (forward's draw is discarded)."""
        out, mu, logvar = self.forward(motion_coef)
        std = torch.exp(0.5 * logvar)
        return mu + torch.randn_like(std) * std

    def __del__(self):
        try:
            if self._h is not None:
                _lib.lib().msmd_style_destroy(self._h)
        except Exception:
            pass
