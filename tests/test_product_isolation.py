"""The product path never routes through the oracle or a CPU fallback: the package sources do not mention oracle/,
importing the package does not import it, and the entry points refuse CPU tensors."""
import ast
import glob
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'ubisoft-laforge-msmd_b200')


def _imports(path):
    mods = set()
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.Import):
            mods.update(a.name.split('.')[0] for a in node.names)
        elif isinstance(node, ast.ImportFrom) and node.module and node.level == 0:
            mods.add(node.module.split('.')[0])
    return mods


def test_package_sources_do_not_import_the_oracle():
    files = glob.glob(os.path.join(PKG, '**', '*.py'), recursive=True) + glob.glob(os.path.join(ROOT, 'msmd_b200', '*.py'))
    assert files
    for f in files:
        assert 'oracle' not in _imports(f), f
    assert 'oracle' not in _imports(os.path.join(ROOT, 'tools', 'synth.py'))
    for f in glob.glob(os.path.join(PKG, 'csrc', '*')):
        assert 'oracle' not in open(f, errors='ignore').read(), f


def test_importing_the_package_does_not_import_the_oracle():
    code = ('import sys; sys.path.insert(0, %r); import msmd_b200, msmd_b200.model, msmd_b200.inference, msmd_b200.decode; '
            'assert not any(m == "oracle" or m.startswith("oracle.") for m in sys.modules), "oracle imported"') % ROOT
    subprocess.run([sys.executable, '-c', code], check=True)


def test_bench_gpu_arm_only_uses_the_oracle_in_its_cpu_baseline_legs():
    src = open(os.path.join(ROOT, 'bench.py')).read()
    tree = ast.parse(src)
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef)]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            uses = any(isinstance(n, (ast.Import, ast.ImportFrom)) and 'oracle' in ast.dump(n) for n in ast.walk(fn))
            assert not uses or fn.name == 'cpu_reference', (cls.name, fn.name)
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'main')
    assert not any(isinstance(n, (ast.Import, ast.ImportFrom)) and 'oracle' in ast.dump(n) for n in ast.walk(main))


def test_cpu_tensors_are_refused():
    from msmd_b200 import _lib
    from msmd_b200.utils import rotation_conversions as R
    with pytest.raises(_lib.MsmdError):
        R.euler_angles_to_matrix(torch.zeros(2, 3), 'YXZ')
