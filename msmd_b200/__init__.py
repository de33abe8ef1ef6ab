"""Import alias: the package directory is ``ubisoft-laforge-msmd_b200`` (not an identifier),
so ``import msmd_b200`` resolves to it by pointing this package's search path there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'ubisoft-laforge-msmd_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
