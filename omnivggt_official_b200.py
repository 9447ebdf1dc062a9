"""Import alias: the package directory is named ``omnivggt-official_b200`` (not a valid Python identifier), so
``import omnivggt_official_b200`` resolves to it through this shim.  No code lives here."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "omnivggt-official_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
