"""Drop-in for the reference's python/device_utils.py:8-13.  The reference returns `cpu` on
Linux even when CUDA is present; this path exists on CUDA only."""
from _load import renderer as _r

get_default_device = _r.get_default_device
