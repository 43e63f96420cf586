"""Entry-point module for redback's plugin loader (`redback.model.modules`, model_library.py:50-87).
Built-ins win on name collisions (model_library.py:66-71), so the batched functions are exported with a
`_b200` suffix."""
from . import all_models_dict as _models

for _name, _fn in _models.items():
    globals()[_name + "_b200"] = _fn
__all__ = [n + "_b200" for n in _models]
