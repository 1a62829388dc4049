"""Import alias: the product package lives in the directory the project layout mandates
(`traffic-context-augmented-vehicle-trajectory-prediction-framework-using-multimodal-llm_b200/`), whose
name is not a Python identifier.  `import tcavp_b200` resolves every submodule from that directory."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "traffic-context-augmented-vehicle-trajectory-prediction-framework-using-multimodal-llm_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
