"""Nav/sig-aware shape tuple: minimal mirror of the reference ``libertem.common.Shape``
(src/libertem/common/shape.py:7-213) -- only what the masked-reduction path touches."""
import math


class Shape(tuple):
    """A shape whose last ``sig_dims`` entries are signal dimensions."""

    def __new__(cls, shape, sig_dims):
        obj = super().__new__(cls, tuple(int(s) for s in shape))
        obj._sig_dims = int(sig_dims)
        return obj

    def __getnewargs__(self):
        return (tuple(self), self._sig_dims)

    @property
    def sig_dims(self):
        return self._sig_dims

    @property
    def nav_dims(self):
        return len(self) - self._sig_dims

    @property
    def dims(self):
        return len(self)

    @property
    def nav(self):
        return Shape(tuple(self)[:self.nav_dims], sig_dims=0)

    @property
    def sig(self):
        return Shape(tuple(self)[self.nav_dims:], sig_dims=self._sig_dims)

    @property
    def size(self):
        return int(math.prod(self)) if len(self) else 0

    def to_tuple(self):
        return tuple(self)

    def flatten_nav(self):
        return Shape((self.nav.size,) + tuple(self.sig), sig_dims=self._sig_dims)

    def flatten_sig(self):
        return Shape(tuple(self.nav) + (self.sig.size,), sig_dims=1)

    def __repr__(self):
        return repr(tuple(self))
