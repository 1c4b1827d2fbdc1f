"""Import shim: lets the reference's `import sparseconvnet as scn` (mopa/models/scn_unet.py:4) resolve to the
B200-native implementation in mopa_b200.scn without editing the reference. Put the repo root on PYTHONPATH."""
from mopa_b200.scn import *  # noqa: F401,F403
from mopa_b200.scn import __all__  # noqa: F401
