"""SpectralArray / SpectralVector / BlockSpectralMatrix: the array types that cross the
boundary, mirroring /root/reference/src/spectralarray.jl:5-98 and
src/blockspectralmatrix.jl:5-129 as far as the hot path and its callers need them.

A SpectralArray is a dense float64 numpy array plus the multipole of its first element
along each axis (an OffsetArray, 0-indexed by default).  `A[l1, l2]` indexes by multipole;
`A.parent` is the dense array.  Matrices are stored column-major (order="F") so that
`parent` has exactly the memory layout of Julia's `parent(::SpectralArray)`.
"""
from __future__ import annotations

import numpy as np


class SpectralArray:
    __slots__ = ("parent", "offsets")

    def __init__(self, parent, offsets=None):
        parent = np.asarray(parent, dtype=np.float64)
        if parent.ndim == 2 and not parent.flags.f_contiguous:
            parent = np.asfortranarray(parent)
        self.parent = parent
        if offsets is None:
            offsets = (0,) * parent.ndim
        if np.isscalar(offsets):
            offsets = (int(offsets),)
        self.offsets = tuple(int(o) for o in offsets)
        if len(self.offsets) != parent.ndim:
            raise ValueError("one offset per axis")

    # -- axes ------------------------------------------------------------------------
    @property
    def ndim(self):
        return self.parent.ndim

    @property
    def shape(self):
        return self.parent.shape

    def axes(self, k=None):
        r = tuple(range(o, o + n) for o, n in zip(self.offsets, self.parent.shape))
        return r if k is None else r[k]

    def firstindex(self, k=0):
        return self.offsets[k]

    def lastindex(self, k=0):
        return self.offsets[k] + self.parent.shape[k] - 1

    # -- multipole indexing ----------------------------------------------------------
    def _key(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        out, new_off = [], []
        for k, o, n in zip(key, self.offsets, self.parent.shape):
            if isinstance(k, slice):
                start = o if k.start is None else k.start
                stop = o + n if k.stop is None else k.stop      # python-style exclusive stop
                if start < o or stop > o + n:
                    raise IndexError("multipole slice outside the array")
                out.append(slice(start - o, stop - o, k.step))
                new_off.append(start)
            else:
                k = int(k)
                if k < o or k >= o + n:
                    raise IndexError(f"multipole {k} outside {o}:{o + n - 1}")
                out.append(k - o)
        return tuple(out), tuple(new_off)

    def __getitem__(self, key):
        k, off = self._key(key)
        r = self.parent[k]
        return SpectralArray(r, off) if off else float(r)

    def __setitem__(self, key, value):
        k, _ = self._key(key)
        self.parent[k] = value

    def __len__(self):
        return self.parent.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.parent, dtype=dtype)

    def __repr__(self):
        ax = ", ".join(f"{a.start}:{a.stop - 1}" for a in self.axes())
        return f"SpectralArray[{ax}]\n{self.parent!r}"

    def copy(self):
        return SpectralArray(self.parent.copy(order="K"), self.offsets)

    def zero_based(self, lmax):
        """Contiguous x[l], l = 0..lmax, as the C ABI wants it (entries below the first
        stored multipole are never read by the loops and are filled with 0)."""
        if self.ndim != 1:
            raise ValueError("zero_based is for vectors")
        o, n = self.offsets[0], self.parent.shape[0]
        if o + n - 1 < lmax:
            raise ValueError(f"vector ends at l={o + n - 1}, need lmax={lmax}")
        out = np.zeros(lmax + 1)
        lo = max(o, 0)
        out[lo:] = self.parent[lo - o:lmax + 1 - o]
        return out

    # -- linear algebra used around the hot path (host LAPACK, as in the reference) ---
    def solve(self, b):
        """`M \\ b` (src/blockspectralmatrix.jl:124-129)."""
        if self.ndim != 2:
            raise ValueError("solve needs a matrix")
        bb = b.parent if isinstance(b, SpectralArray) else np.asarray(b, dtype=np.float64)
        if isinstance(b, SpectralArray) and b.offsets[0] != self.offsets[1]:
            raise ValueError("first multipole of the right-hand side must match the matrix columns")
        x = np.linalg.solve(self.parent, bb)
        return SpectralArray(x, (self.offsets[1],) + ((b.offsets[1],) if getattr(b, "ndim", 1) == 2 else ()))

    def inv(self):
        return SpectralArray(np.linalg.inv(self.parent), self.offsets)


def SpectralVector(a, offset=0):
    """0-indexed vector by default (src/spectralarray.jl:16)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("SpectralVector needs a 1-d array")
    return SpectralArray(a, (offset,))


def _rng(r):
    if isinstance(r, range):
        return r.start, len(r)
    if isinstance(r, int):
        return 0, r
    raise TypeError("sizes or ranges")


def spectralzeros(*ranges):
    """spectralzeros(range1, range2, ...) (src/spectralarray.jl:73-74)."""
    offs, shape = zip(*[_rng(r) for r in ranges])
    return SpectralArray(np.zeros(shape, order="F"), offs)


_PINNED = {}          # address of the first element -> the library block that owns the memory


def pinned_spectralzeros(r, interleave=False):
    """A zero N x N SpectralArray over the multipoles of `r` (a range or a size) in page-locked memory of the library
    (psb200_host_alloc; `pinned_spectralzeros` of julia/PowerSpectraB200.jl): the result of a host call lands in it at the
    full PCIe rate.  `interleave=True` spreads it over the NUMA nodes of the host (calls on several GPUs).  Pass it to
    the in-place functions (inner_mcm00, loop_covTTTT, ...); release it with `free_pinned` -- the memory is not
    garbage collected, and the array must not be touched afterwards."""
    from ._lib import HostMatrix
    lo, n = _rng(r)
    H = HostMatrix(n, interleave)
    A = SpectralArray(H.array, (lo, lo))
    _PINNED[A.parent.ctypes.data] = H
    return A


def free_pinned(A):
    H = _PINNED.pop(A.parent.ctypes.data, None)
    if H is None:
        raise ValueError("not a pinned_spectralzeros array")
    A.parent = None
    H.free()


def spectralones(*ranges):
    offs, shape = zip(*[_rng(r) for r in ranges])
    return SpectralArray(np.ones(shape, order="F"), offs)


class BlockSpectralMatrix:
    """hvcat of SpectralArrays into one dense matrix that remembers the multipole range of
    every block (src/blockspectralmatrix.jl:5-42)."""

    def __init__(self, blocks):
        rows = [list(r) for r in blocks]
        self.parent = np.asfortranarray(np.block([[b.parent for b in r] for r in rows]))
        self.m_ells = tuple(r[0].axes(0) for r in rows)
        self.n_ells = tuple(b.axes(1) for b in rows[0])

    def solve(self, b):
        """`A \\ [x; y]` for stacked vectors (src/blockspectralmatrix.jl:100-122)."""
        if isinstance(b, (list, tuple)):
            rhs = np.concatenate([v.parent for v in b])
        else:
            rhs = np.asarray(b, dtype=np.float64)
        x = np.linalg.solve(self.parent, rhs)
        out, o = [], 0
        for r in self.n_ells:
            out.append(SpectralArray(x[o:o + len(r)], (r.start,)))
            o += len(r)
        return tuple(out)

    def getblock(self, i, j):
        ro = sum(len(r) for r in self.m_ells[:i])
        co = sum(len(r) for r in self.n_ells[:j])
        return SpectralArray(self.parent[ro:ro + len(self.m_ells[i]), co:co + len(self.n_ells[j])],
                             (self.m_ells[i].start, self.n_ells[j].start))


def decouple_covmat(Y: SpectralArray, B1: SpectralArray, B2: SpectralArray) -> SpectralArray:
    """B1^-1 Y (B2^-1)^T on the host (src/covariance.jl:8-14), as the reference does it (LAPACK)."""
    C = np.linalg.solve(B1.parent, Y.parent)
    C = np.linalg.solve(B2.parent, C.T).T
    return SpectralArray(C, Y.offsets)


def decouple_covmat_device(Y: SpectralArray, B1: SpectralArray, B2: SpectralArray) -> SpectralArray:
    """The same on the GPU: both LU factorisations (of B1', B2' like the reference) and the two
    n-right-hand-side solves run in libpsb200 (psb200_decouple_covmat: cuSOLVER getrf/getrs on the current device)."""
    from . import _lib
    n = Y.parent.shape[0]
    if Y.parent.shape != (n, n) or B1.parent.shape != (n, n) or B2.parent.shape != (n, n):
        raise ValueError("decouple_covmat needs three square matrices of one size")
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    y, b1, b2 = f(Y.parent), f(B1.parent), f(B2.parent)
    out = np.zeros((n, n), order="F")
    dp = lambda a: a.ctypes.data_as(_lib.DP)
    _lib.check(_lib.lib().psb200_decouple_covmat(n, dp(y), n, dp(b1), n, dp(b2), n, dp(out), n))
    return SpectralArray(out, Y.offsets)
