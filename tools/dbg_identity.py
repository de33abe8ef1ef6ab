import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import test_parallel as tp
from helpers import make_msmd
dev = torch.device('cuda', 0)
import msmd_b200.model as M
for prec in ('bf16', 'fp16', 'fp32', 'hybrid'):
    for noise in ('z', 'philox'):
        orig = tp.__dict__['_generate_real']
        def gen(lo, hi):
            import helpers
            mm = helpers.make_msmd
            helpers.make_msmd = lambda d, precision=None, **kw: mm(d, precision=prec, **kw)
            try:
                return orig(lo, hi, dev, noise)
            finally:
                helpers.make_msmd = mm
        whole = gen(0, 6); part = gen(2, 6)
        d = (part - whole[2:6]).abs().amax(dim=(1, 2))
        print(prec, noise, 'max abs diff per clip', d.tolist(), 'first differing frame', (part - whole[2:6]).abs().amax(dim=(0, 2)).nonzero().flatten()[:3].tolist())
