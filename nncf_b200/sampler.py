"""Drop-in for the reference's Cython class `sampler.MultinomialSampler`.

ref: sampler/sampler.pyx:19-38 — MultinomialSampler(np.float64[n] dist, int dist_size, double neg_sampling_power=0.75,
unsigned long long rand_seed=0); .sample() -> int; .sample_batch(int n) -> np.ndarray[int32, n].
Same names, argument meaning and dtypes; the draws come from the device alias-table + Philox sampler.  Deviation
(declared): the reference seeds from time(NULL) when rand_seed == 0 and reads an uninitialised seed otherwise
(sampler/nodesampler.cpp:59-60), so its streams are not reproducible; here rand_seed is the Philox key and the stream
is fully determined by it.  `sample_batch_device` is the extra entry point that writes into device memory.
"""
from __future__ import annotations

import numpy as np

from . import ops


class MultinomialSampler(object):
    """sample from given categorical distribution.  Note: sample batch size: int, sample node index: int"""

    def __init__(self, input, dist_size, neg_sampling_power=0.75, rand_seed=0):
        if input is None:
            raise TypeError("Argument 'input' must not be None")
        input = np.ascontiguousarray(input, dtype=np.float64)
        if input.ndim != 1:
            raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % input.ndim)
        dist_size = int(dist_size)
        assert dist_size <= input.size
        self._dev = ops.DeviceSampler(input[:dist_size], float(neg_sampling_power), int(rand_seed))
        self._buf = np.zeros(0, dtype=np.int32)
        self._pos = 0

    def sample(self):
        """return a node index (int).  Single draws are served from a prefetched block of the same stream."""
        if self._pos >= self._buf.size:
            self._buf = self._dev.sample_host(4096)
            self._pos = 0
        v = int(self._buf[self._pos])
        self._pos += 1
        return v

    def sample_batch(self, n):
        """return an array of node index (np.int32[n])"""
        return self._dev.sample_host(int(n))

    def sample_batch_device(self, n, out=None):
        """n draws into a CUDA int32 tensor (no host round trip)"""
        return self._dev.sample_device(int(n), out)
