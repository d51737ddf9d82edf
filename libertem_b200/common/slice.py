"""Origin+shape slices: minimal mirror of ``libertem.common.Slice``
(src/libertem/common/slice.py:17-400) for the operations the hot path needs
(sig slicing of mask stacks, tile/partition addressing, mask shifts)."""
from .shape import Shape


class Slice:
    __slots__ = ('origin', 'shape')

    def __init__(self, origin, shape):
        if not isinstance(shape, Shape):
            raise TypeError('shape must be a Shape')
        origin = tuple(int(o) for o in origin)
        if len(origin) != len(shape):
            raise ValueError('origin and shape need the same number of dimensions')
        self.origin = origin
        self.shape = shape

    def __repr__(self):
        return f'<Slice origin={self.origin} shape={self.shape}>'

    def __eq__(self, other):
        return (isinstance(other, Slice) and self.origin == other.origin
                and tuple(self.shape) == tuple(other.shape)
                and self.shape.sig_dims == other.shape.sig_dims)

    def __hash__(self):
        return hash((self.origin, tuple(self.shape), self.shape.sig_dims))

    def is_null(self):
        return any(s <= 0 for s in self.shape)

    def get(self, arr=None, sig_only=False, nav_only=False):
        """tuple of python slices (or ``arr`` indexed with it) -- slice.py:167-226."""
        o, s = self.origin, tuple(self.shape)
        nd = self.shape.nav_dims
        if sig_only and nav_only:
            raise ValueError('sig_only and nav_only are mutually exclusive')
        if sig_only:
            o, s = o[nd:], s[nd:]
        elif nav_only:
            o, s = o[:nd], s[:nd]
        sl = tuple(slice(a, a + b) for a, b in zip(o, s))
        if arr is None:
            return sl
        if sig_only:
            return arr[(Ellipsis,) + sl]
        return arr[sl]

    def discard_nav(self):
        nd = self.shape.nav_dims
        return Slice(origin=self.origin[nd:], shape=self.shape.sig)

    @property
    def sig(self):
        return self.discard_nav()

    @property
    def nav(self):
        nd = self.shape.nav_dims
        return Slice(origin=self.origin[:nd], shape=self.shape.nav)

    def shift_by(self, offsets):
        """Shift the origin by ``offsets`` (slice.py, used by masks.py:85-124)."""
        offsets = tuple(int(v) for v in offsets)
        if len(offsets) != len(self.origin):
            raise ValueError('shift must match the number of dimensions')
        return Slice(origin=tuple(a + b for a, b in zip(self.origin, offsets)), shape=self.shape)

    def intersection_with(self, other):
        lo = tuple(max(a, b) for a, b in zip(self.origin, other.origin))
        hi = tuple(min(a + s, b + t) for a, s, b, t in
                   zip(self.origin, self.shape, other.origin, other.shape))
        shp = tuple(max(0, h - l) for l, h in zip(lo, hi))
        return Slice(origin=lo, shape=Shape(shp, sig_dims=self.shape.sig_dims))
