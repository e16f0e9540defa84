"""Import alias: the package directory is `llm-awq_b200/` (not a valid Python identifier),
so `import llm_awq_b200` executes that directory's __init__ under this name."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "llm-awq_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _os, _f
