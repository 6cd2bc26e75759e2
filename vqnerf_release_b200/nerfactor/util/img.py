"""Mirror of decomp/nerfvq_nfr3/nerfactor/util/img.py:142-186."""
from ...abi import linear2srgb, srgb2linear  # noqa: F401
