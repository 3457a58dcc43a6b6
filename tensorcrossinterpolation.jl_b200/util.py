"""Host-side helpers mirroring src/util.jl, src/sweepstrategies.jl and the index-set
conventions of src/tensorci2.jl.  Index sets are int64 arrays of shape (count, len)
(one multi-index per row, 1-based values), the array form of Vector{MultiIndex}."""
import numpy as np

M64 = (1 << 64) - 1


def jl_max(x, y):
    """Julia's max for Float64: NaN-propagating (util.jl:1-10 relies on it)."""
    if x != x or y != y:
        return float("nan")
    return x if x > y else y


def forwardsweep(sweepstrategy, iteration):  # sweepstrategies.jl:1-6
    return sweepstrategy == "forward" or (sweepstrategy == "backandforth" and iteration % 2 == 1)


def as_indexset(x, length=None):
    """Vector{MultiIndex} -> (count, len) int64 array."""
    if isinstance(x, np.ndarray) and x.ndim == 2:
        return np.ascontiguousarray(x, dtype=np.int64)
    n = len(x)
    if n == 0:
        return np.zeros((0, length or 0), dtype=np.int64)
    ln = len(x[0]) if length is None else length
    return np.ascontiguousarray(np.asarray(x, dtype=np.int64).reshape(n, ln))


def kronecker_left(Iset, localdim):  # kronecker(Iset, d)  tensorci2.jl:315-320: i fastest, then sigma
    n, l = Iset.shape
    out = np.empty((n * localdim, l + 1), dtype=np.int64)
    out[:, :l] = np.tile(Iset, (localdim, 1))
    out[:, l] = np.repeat(np.arange(1, localdim + 1, dtype=np.int64), n)
    return out


def kronecker_right(localdim, Jset):  # kronecker(d, Jset)  tensorci2.jl:322-327: sigma fastest, then j
    n, l = Jset.shape
    out = np.empty((n * localdim, l + 1), dtype=np.int64)
    out[:, 1:] = np.repeat(Jset, localdim, axis=0)
    out[:, 0] = np.tile(np.arange(1, localdim + 1, dtype=np.int64), n)
    return out


def _rowkeys(a):
    """One hashable key per row (the raw bytes of the row)."""
    a = np.ascontiguousarray(a)
    return a.view(np.dtype((np.void, 8 * a.shape[1]))).ravel().tolist()


def union(a, b):
    """Base.union(a, b) on vectors of multi-indices: order preserving, duplicates dropped.  `a` is a
    kronecker product of a duplicate-free set (tensorci2.jl:526-527), hence already duplicate free."""
    if b.shape[0] == 0:
        return a
    if a.shape[1] == 0:
        return a[:1] if a.shape[0] else b[:1]
    seen = set(_rowkeys(a))
    keep = []
    for i, k in enumerate(_rowkeys(b)):
        if k not in seen:
            seen.add(k)
            keep.append(i)
    if not keep:
        return a
    return np.concatenate([a, b[keep]], axis=0)


def pushunique(arr, item):  # util.jl:16-20 on an index-set array
    item = np.asarray(item, dtype=np.int64).reshape(1, -1)
    if arr.shape[0] and (arr == item).all(axis=1).any():
        return arr
    return np.concatenate([arr.reshape(arr.shape[0], item.shape[1]), item], axis=0)


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def _splitmix64_np(x):  # the same on uint64 arrays (wrapping arithmetic)
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def uniform01(seed, index):  # tci_uniform01 of include/tci_targets.h
    h = _splitmix64(_splitmix64(seed & M64) ^ _splitmix64((index + 0x632BE59BD9B4E019) & M64))
    return (h >> 11) * (1.0 / 9007199254740992.0)


def pushrandomsubset(subset, N, cnt, rng):
    """pushrandomsubset!(subset, 1:N, cnt) (util.jl:36-58); subset is a list of 1-based ints."""
    have = set(subset)
    c = [v for v in range(1, N + 1) if v not in have]
    for _ in range(min(cnt, len(c))):
        index = rng.randindex(len(c))
        subset.append(c.pop(index - 1))


class CounterRNG:
    """Injected replacement for `rng` in the global pivot finder (globalpivotfinder.jl:156): the
    reference draws start points from Julia's Xoshiro, which cannot be reproduced here, so the
    host layer and the oracle share this counter-based generator instead."""

    def __init__(self, seed=1):
        self.seed = int(seed)
        self.calls = 0
        self.rook_draws = 0

    def randindex(self, length):
        """rand(1:length) of randomsubset (util.jl:47), shared with the oracle's RookRng."""
        v = 1 + int(uniform01(self.seed ^ 0x726F6F6B, self.rook_draws) * float(length))
        self.rook_draws += 1
        return min(v, int(length))

    def start_points(self, nsearch, localdims):
        self.calls += 1
        n = len(localdims)
        ld = np.asarray(localdims, dtype=np.int64)
        with np.errstate(over="ignore"):  # uint64 arithmetic wraps, as the C generator's does
            s = np.arange(nsearch, dtype=np.uint64)[:, None]
            p = np.arange(n, dtype=np.uint64)[None, :]
            idx = (np.uint64(self.calls * 1000003 & M64) + s) * np.uint64(1009) + p
            h = _splitmix64_np(np.uint64(_splitmix64(self.seed & M64)) ^ _splitmix64_np(idx + np.uint64(0x632BE59BD9B4E019)))
        u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return np.minimum(1 + (u * ld[None, :].astype(np.float64)).astype(np.int64), ld[None, :])
