"""CPU: the C-ABI library loads and exports every symbol include/msmd_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'msmd_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(msmd_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_entry_points():
    fns = header_functions()
    for must in ('msmd_rot_convert', 'msmd_flame_create', 'msmd_flame_decode', 'msmd_last_error'):
        assert must in fns


def test_library_exports_all_declared_symbols(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [f for f in header_functions() if not hasattr(lib, f)]
    assert not missing, f'declared in the header but not exported: {missing}'


def test_python_binding_covers_header(built_lib):
    from msmd_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    assert _lib.lib().msmd_version().decode().endswith('sm_100a')


def test_argument_errors_without_gpu(built_lib):
    """Pure argument validation never touches the device."""
    from msmd_b200 import _lib
    l = _lib.lib()
    assert l.msmd_rot_convert(999, ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 0, None) == -1
    assert b'unknown kind' in l.msmd_last_error()
    assert l.msmd_rot_convert(2, ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 0 * 9 + 0 * 3 + 1, None) == -1  # "XXY"
    assert l.msmd_rot_convert(0, None, None, 0, 0, None) == 0      # empty input is a no-op
    assert l.msmd_flame_decode(None, None, None, 1, 4, None, None, 0, None) == -1


def test_no_cpu_fallback(built_lib):
    import pytest
    import torch
    from msmd_b200 import _lib
    from msmd_b200.utils import rotation_conversions as rc
    with pytest.raises(_lib.MsmdError):
        rc.axis_angle_to_matrix(torch.zeros(4, 3))


def test_convention_errors_match_reference(built_lib):
    import pytest
    import torch
    from msmd_b200.utils import rotation_conversions as rc
    x = torch.zeros(2, 3)
    for bad, msg in (('XY', 'Convention must have 3 letters.'), ('XXY', 'Invalid convention XXY.'),
                     ('XAZ', 'Invalid letter A in convention string.')):
        with pytest.raises(ValueError, match=msg):
            rc.euler_angles_to_matrix(x, bad)
        with pytest.raises(ValueError, match=msg):
            rc.matrix_to_euler_angles(torch.zeros(2, 3, 3), bad)
    with pytest.raises(ValueError, match='Invalid input euler angles.'):
        rc.euler_angles_to_matrix(torch.zeros(2, 4), 'XYZ')
    with pytest.raises(ValueError, match='Invalid rotation matrix'):
        rc.matrix_to_quaternion(torch.zeros(2, 3, 4))
    with pytest.raises(ValueError, match='Points are not in 3D'):
        rc.quaternion_apply(torch.zeros(2, 4), torch.zeros(2, 4))


def test_plain_c_consumer_compiles_links_and_runs(built_lib, tmp_path):
    """include/msmd_b200.h is a C header (extern "C", plain pointers and sizes): a C99 translation unit binds the
    library directly, which is what a cgo / JNI / N-API stub of the reference would do."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    assert gcc, 'gcc is part of the image'
    exe = str(tmp_path / 'c_abi_consumer')
    libdir = os.path.dirname(built_lib)
    subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                    os.path.join(ROOT, 'tests', 'c_abi_consumer.c'), '-o', exe, '-L', libdir, '-lmsmd_b200',
                    '-Wl,-rpath,' + libdir, '-Wl,-rpath,/usr/local/cuda/lib64'], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith('ok '), (r.returncode, r.stdout, r.stderr)
