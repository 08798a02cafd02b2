"""Import alias: the package directory is `h-denseunet_b200/` (not a valid identifier), so this
shim points the importable name `h_denseunet_b200` at it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "h-denseunet_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
