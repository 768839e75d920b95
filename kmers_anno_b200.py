"""Import shim: the package directory is `kmers.anno_b200/` (the dot is part of the
reference's name and cannot appear in a Python identifier), so `import kmers_anno_b200`
loads that directory as a regular package under this name."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "kmers.anno_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
