"""Minimal stand-in for pydata ``sparse`` (third-party, absent in this image).

Only what the reference's MaskContainer / mask generators need to *stack,
slice and hand over* COO masks to scipy.sparse: no arithmetic on the hot path
happens here (that is scipy CSR + the reference's numba ``rmatmul``).
See tests/golden/shims/sparseconverter/__init__.py for why this exists.
"""
import numpy as np


class SparseArray:
    pass


class COO(SparseArray):
    def __init__(self, coords=None, data=None, shape=None, fill_value=0, **kw):
        if coords is not None and data is None and shape is None and not isinstance(coords, tuple):
            other = coords
            if isinstance(other, COO):
                coords, data, shape = other.coords, other.data, other.shape
            else:
                o = COO.from_numpy(np.asarray(other))
                coords, data, shape = o.coords, o.data, o.shape
        coords = np.asarray(coords, dtype=np.int64)
        if coords.ndim == 1:
            coords = coords[None, :]
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        data = np.asarray(data)
        if data.ndim == 0:
            data = np.full(coords.shape[1], data[()])
        self.shape = tuple(int(s) for s in shape)
        # canonical: sorted lexicographically, duplicates summed
        if coords.shape[1]:
            lin = np.ravel_multi_index(tuple(coords), self.shape)
            order = np.argsort(lin, kind='stable')
            lin = lin[order]
            data = data[order]
            uniq, start = np.unique(lin, return_index=True)
            if len(uniq) != len(lin):
                data = np.add.reduceat(data, start)
                lin = uniq
            coords = np.stack(np.unravel_index(lin, self.shape)).astype(np.int64)
        self.coords = coords.reshape((len(self.shape), -1))
        self.data = data
        self.fill_value = fill_value

    dtype = property(lambda self: self.data.dtype)
    ndim = property(lambda self: len(self.shape))
    nnz = property(lambda self: self.data.shape[0])

    @classmethod
    def from_numpy(cls, a):
        a = np.asarray(a)
        nz = np.nonzero(a)
        return cls(coords=np.stack(nz) if a.ndim else np.zeros((0, 0)), data=a[nz], shape=a.shape)

    @classmethod
    def from_scipy_sparse(cls, m):
        m = m.tocoo()
        return cls(coords=np.stack((m.row, m.col)), data=m.data, shape=m.shape)

    def todense(self):
        out = np.zeros(self.shape, dtype=self.data.dtype)
        out[tuple(self.coords)] = self.data
        return out

    def __array__(self, dtype=None, copy=None):
        d = self.todense()
        return d if dtype is None else d.astype(dtype)

    def astype(self, dtype, **kw):
        return COO(coords=self.coords, data=self.data.astype(dtype), shape=self.shape)

    def reshape(self, shape):
        if isinstance(shape, (int, np.integer)):
            shape = (shape,)
        shape = list(shape)
        size = int(np.prod(self.shape))
        if -1 in shape:
            i = shape.index(-1)
            rest = int(np.prod([s for s in shape if s != -1]))
            shape[i] = size // rest if rest else 0
        lin = np.ravel_multi_index(tuple(self.coords), self.shape) if self.nnz else \
            np.zeros(0, dtype=np.int64)
        coords = np.stack(np.unravel_index(lin, shape)) if len(shape) else np.zeros((0, 0))
        return COO(coords=coords, data=self.data, shape=tuple(shape))

    @property
    def T(self):
        return COO(coords=self.coords[::-1], data=self.data, shape=self.shape[::-1])

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if Ellipsis in key:
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        sel = np.ones(self.nnz, dtype=bool)
        new_coords = []
        new_shape = []
        for ax, k in enumerate(key):
            c = self.coords[ax]
            if isinstance(k, (int, np.integer)):
                kk = int(k) % self.shape[ax]
                sel &= (c == kk)
            else:
                start, stop, step = k.indices(self.shape[ax])
                assert step == 1
                sel &= (c >= start) & (c < stop)
                new_coords.append((ax, start))
                new_shape.append(max(0, stop - start))
        if not new_shape:
            v = self.data[sel]
            return v[0] if len(v) else self.data.dtype.type(0)
        coords = np.stack([self.coords[ax][sel] - start for ax, start in new_coords])
        return COO(coords=coords, data=self.data[sel], shape=tuple(new_shape))

    def _binary(self, other, op):
        if isinstance(other, COO):
            return COO.from_numpy(op(self.todense(), other.todense()))
        return COO.from_numpy(op(self.todense(), other))

    def __add__(self, other):
        return self._binary(other, np.add)

    def __mul__(self, other):
        return self._binary(other, np.multiply)

    __rmul__ = __mul__

    def sum(self, *a, **kw):
        return self.todense().sum(*a, **kw)


class GCXS(COO):
    pass


class DOK(COO):
    pass


def concatenate(arrays, axis=0):
    arrays = list(arrays)
    coords, data, off = [], [], 0
    for a in arrays:
        c = a.coords.copy()
        c[axis] += off
        off += a.shape[axis]
        coords.append(c)
        data.append(a.data)
    dt = np.result_type(*[d.dtype for d in data])
    shape = list(arrays[0].shape)
    shape[axis] = off
    return COO(coords=np.concatenate(coords, axis=1),
               data=np.concatenate([d.astype(dt) for d in data]),
               shape=tuple(shape))


def stack(arrays, axis=0):
    assert axis == 0
    return concatenate([a.reshape((1,) + a.shape) for a in arrays])
