"""Is the reference package (`torchebm`) importable next to this one?

When it is, `torchebm_b200` builds its samplers, integrators, loss and schedulers ON TOP of the reference's own
classes (`torchebm_b200/dropin.py`): `torchebm_b200.LangevinDynamics` is then a subclass of
`torchebm.samplers.LangevinDynamics` that takes the fused CUDA path when it applies and calls the reference's own
`sample()` (`super()`) for everything else -- other dtypes, CPU devices, conditioning, custom integrators.  When it is
not (e.g. a GPU box without the reference), the package's standalone mirrors of that API surface are used
(`samplers.py`, `integrators.py`, `losses.py`, `core.py`), which have no such delegate and raise instead.

`EBM_B200_STANDALONE=1` forces the standalone classes even when `torchebm` is importable.
"""

from __future__ import annotations

import os
from types import ModuleType
from typing import Optional

_cached: Optional[ModuleType] = None
_probed = False


def reference() -> Optional[ModuleType]:
    global _cached, _probed
    if _probed:
        return _cached
    _probed = True
    if os.environ.get("EBM_B200_STANDALONE", "") not in ("", "0"):
        return None
    try:
        import torchebm  # noqa: F401
        import torchebm.core  # noqa: F401
        import torchebm.samplers  # noqa: F401
        import torchebm.integrators  # noqa: F401
        import torchebm.losses  # noqa: F401
    except Exception:  # noqa: BLE001  (missing, or broken by its own optional dependencies)
        return None
    _cached = torchebm
    return _cached
