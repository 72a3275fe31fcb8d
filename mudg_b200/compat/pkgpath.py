"""Let the drop-in packages (`lvdm`, `utils`, `virtual_render`) coexist with a checkout of the reference on sys.path.

The reference's packages are namespace packages (no __init__.py); this repo's are regular packages and therefore shadow
them completely.  Each drop-in package extends its __path__ with the same-named directories found LATER on sys.path, so
modules this repo does not replace -- the driver `virtual_render/virtual_pose_render.py`, the dataset readers
`virtual_render/data_tools.py`, the OpenCLIP embedders `lvdm/modules/encoders/condition.py` -- still resolve to the
reference, while every module that exists here wins (its directory stays first in __path__)."""
from __future__ import annotations

import os
import sys
from typing import List


def extended(path: List[str], name: str) -> List[str]:
    rel = name.replace(".", os.sep)
    out = [os.path.abspath(p) for p in path]
    for entry in sys.path:
        d = os.path.abspath(os.path.join(entry or os.getcwd(), rel))
        if os.path.isdir(d) and d not in out:
            out.append(d)
    return out
