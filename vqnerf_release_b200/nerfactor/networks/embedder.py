"""Mirror of decomp/nerfvq_nfr3/nerfactor/networks/embedder.py:23-47."""
import math

from ... import abi


class Embedder:
    def __init__(self, incl_input=True, in_dims=3, log2_max_freq=3, n_freqs=4, log_sampling=True,
                 periodic_func=None):
        # The kernel implements exactly the configuration models/shape.py:82-89 builds.
        if not incl_input or in_dims != 3 or not log_sampling or periodic_func is not None:
            raise NotImplementedError('only incl_input=True, in_dims=3, log_sampling=True, [sin, cos] is built')
        if n_freqs > 1 and not math.isclose(log2_max_freq, n_freqs - 1):
            raise NotImplementedError('freq bands must be 2**linspace(0, n_freqs-1, n_freqs)')
        self.n_freqs = int(n_freqs)
        self.out_dims = 3 + 6 * self.n_freqs

    def __call__(self, x):
        return abi.embed(x, self.n_freqs)
