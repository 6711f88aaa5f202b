"""Import alias: `import show_and_tell_b200` loads the package that lives in `show-and-tell_b200/`
(a directory name Python cannot import directly)."""
import os as _os

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "show-and-tell_b200")
__path__ = [_dir]
__package__ = "show_and_tell_b200"
if globals().get("__spec__") is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(_dir, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
