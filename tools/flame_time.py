"""Per-launch time of the FLAME kernels at config 2 (B = 8192), eager CUDA events: python tools/flame_time.py"""
import sys; sys.path.insert(0,'/root/repo')
import torch
from bench import FlameWorkload
from msmd_b200 import _lib
wl = FlameWorkload(); wl.setup(torch.device('cuda',0), 0)
for _ in range(5): wl.step()
torch.cuda.synchronize()
_lib.lib().msmd_profile_reset(); _lib.lib().msmd_profile_enable(1)
for _ in range(20): wl.step()
torch.cuda.synchronize(); _lib.lib().msmd_profile_enable(0)
print({k: round(v[0]/v[1]*1000,1) for k,v in _lib.profile_dump().items() if k})
