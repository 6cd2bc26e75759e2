"""Mirror of decomp/nerfvq_nfr3/brdf/renderer.py:184-219 (only gen_light_xyz is on the hot path)."""
from ..abi import gen_light_xyz  # noqa: F401  (host float64, computed by the C library)
