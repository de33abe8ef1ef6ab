import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from msmd_b200.utils import rotation_conversions as rc
from oracle import rotations as R
g = torch.Generator().manual_seed(5)
for n in (1, 255, 257, 100_003):
    aa = torch.randn(n, 3, generator=g)
    e = torch.randn(n, 3, generator=g)
    m_ref = R.euler_angles_to_matrix(e.numpy(), 'YXZ')
    m = rc.euler_angles_to_matrix(e.cuda(), 'YXZ').cpu().numpy()
    want = R.matrix_to_axis_angle(m_ref)
    got = rc.euler_angles_to_axis_angle(e.cuda(), 'YXZ').cpu().numpy()
    got2 = rc.matrix_to_axis_angle(torch.from_numpy(m_ref).cuda()).cpu().numpy()
    q_ref = R.matrix_to_quaternion(m_ref)
    q = rc.matrix_to_quaternion(torch.from_numpy(m_ref).cuda()).cpu().numpy()
    err = np.abs(got - want).max(-1)
    i = int(err.argmax())
    print(n, 'matrix err', np.abs(m - m_ref).max(), 'quat err', np.abs(q - q_ref).max(), 'fused aa: frac>5e-5', (err >= 5e-5).mean(), 'max', err.max(),
          '| matrix->aa alone: frac', (np.abs(got2 - want).max(-1) >= 5e-5).mean(), 'max', np.abs(got2 - want).max())
    print('   worst: euler', e[i].numpy(), 'want', want[i], 'got', got[i], 'angle', np.linalg.norm(want[i]))
