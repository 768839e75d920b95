"""Import shim: the package directory is `kmers.anno_b200/` (the dot is part of the
reference's name and cannot appear in a Python identifier), so `import kmers_anno_b200`
resolves here and forwards to that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "kmers.anno_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
__package__ = __name__
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
