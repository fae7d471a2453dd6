"""Import alias: the package directory is `linr-pcgc_b200/` (not a valid identifier), so this
module turns itself into that package.  `import linr_pcgc_b200` == the code under linr-pcgc_b200/."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "linr-pcgc_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
