import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msmd_b200.utils import rotation_conversions as rc
n = 16_000_000
e = torch.randn(n, 3, device='cuda')
for _ in range(3):
    a = rc.euler_angles_to_axis_angle(e, 'YXZ'); m = rc.euler_angles_to_matrix(e, 'YXZ'); b = rc.matrix_to_axis_angle(m)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, byts in (('euler->aa fused', lambda: rc.euler_angles_to_axis_angle(e, 'YXZ'), 24), ('euler->matrix', lambda: rc.euler_angles_to_matrix(e, 'YXZ'), 48),
                       ('matrix->aa', lambda: rc.matrix_to_axis_angle(m), 48)):
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{name}: {n / ms / 1e6:.1f} G rot/s, {n * byts / ms / 1e6:.0f} GB/s (incl. torch.empty + launch)')
