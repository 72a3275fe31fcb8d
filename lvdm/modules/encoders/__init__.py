# drop-in package: modules not replaced here resolve to a reference checkout later on sys.path (mudg_b200/compat/pkgpath.py)
from mudg_b200.compat.pkgpath import extended as _extended

__path__ = _extended(__path__, __name__)
