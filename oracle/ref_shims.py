"""Import the UNMODIFIED reference from /root/reference (build container only).

Used by tests/test_oracle_vs_reference.py and oracle/make_golden.py to pin the
oracle.  The GPU box has no /root/reference: ``available()`` is False there and
callers skip.  The shims are the monkey-patches of SURVEY.md App. B; no
reference file is copied or edited.
"""
import argparse
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get('MSMD_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF, 'model.py'))


from tools.synth import pinned_args  # noqa: E402,F401  (SURVEY App. A configuration; shared with bench.py)


_ready = False


def setup():
    """sys.path + HF from_pretrained + CPU mask shims (App. B items 1-3)."""
    global _ready
    if _ready:
        return
    assert available(), 'reference not present'
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # A package of ours may already be imported as 'utils'/'model'; the reference needs its own.
    for m in ('utils', 'model', 'style_encoder'):
        if m in sys.modules and not getattr(sys.modules[m], '__file__', '').startswith(REF):
            del sys.modules[m]
    import transformers
    from transformers import HubertConfig, Wav2Vec2Config

    def _hub(cls, *a, **k):
        return cls(HubertConfig(attn_implementation='eager'))

    def _w2v(cls, *a, **k):
        return cls(Wav2Vec2Config(attn_implementation='eager'))

    transformers.HubertModel.from_pretrained = classmethod(_hub)
    transformers.Wav2Vec2Model.from_pretrained = classmethod(_w2v)
    import utils.model_common as mc
    import model as ref_model
    if not torch.cuda.is_available():
        ref_model.enc_dec_mask = lambda T, S, fw=2, ex=0, device='cpu': mc.enc_dec_mask(T, S, fw, ex, device='cpu')
    _ready = True


def ref_modules():
    setup()
    import model as ref_model
    import style_encoder as ref_style
    import utils.lbs as ref_lbs
    import utils.rotation_conversions as ref_rc
    import utils.flame as ref_flame
    import utils.model_common as ref_mc
    return types.SimpleNamespace(model=ref_model, style=ref_style, lbs=ref_lbs, rc=ref_rc,
                                 flame=ref_flame, mc=ref_mc)


def ref_infer_coeffs():
    """inference.infer_coeffs with librosa/models/datasets stubbed (App. B item 4)."""
    setup()
    import model as ref_model
    for name in ('librosa', 'datasets', 'models'):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.get_dataset = lambda *a, **k: None
            stub.get_diffusion_model = ref_model.get_diffusion_model
            sys.modules[name] = stub
    import inference
    return inference.infer_coeffs


def ref_flame(raw, n_shape, n_exp):
    """Build the reference FLAME module from a synthetic raw dict (App. B item 7)."""
    m = ref_modules()
    d = tempfile.mkdtemp()
    pkl = os.path.join(d, 'generic_model.pkl')
    with open(pkl, 'wb') as f:
        pickle.dump(raw, f)
    from . import synth
    emb = synth.flame_lmk_embeddings(raw['f'].shape[0])
    npy = os.path.join(d, 'landmark_embedding.npy')
    np.save(npy, emb, allow_pickle=True)
    cfg = types.SimpleNamespace(flame_model_path=pkl, n_shape=n_shape, n_exp=n_exp,
                                flame_lmk_embedding_path=npy)
    return m.flame.FLAME(cfg).eval()


class inject_randn_like:
    """Patch torch.randn_like with supplied tensors, in call order (App. B item 5)."""

    def __init__(self, tensors):
        self.it = iter(tensors)

    def __enter__(self):
        self._orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: next(self.it).to(x.dtype).reshape(x.shape)
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig
