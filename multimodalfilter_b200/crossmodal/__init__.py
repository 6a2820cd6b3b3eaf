"""Mirror of the reference's ``crossmodal`` package layout for the filtering hot path:
``base_models`` (fusion), ``push_models`` / ``door_models`` (task models).  Tasks / datasets /
training helpers are out of scope."""
from . import base_models, door_models, models, push_models  # noqa: F401
