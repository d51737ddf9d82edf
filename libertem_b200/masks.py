"""Mask generators with the reference's signatures (``libertem.masks``,
src/libertem/masks.py): virtual-detector shapes used as inputs of the hot path.

Host-side numpy: masks are built once per run (<= a few MiB) and uploaded; the per-frame work
happens in the CUDA kernels.
"""
import numpy as np


def _disk(centerX, centerY, imageSizeX, imageSizeY, radius):
    ys = np.arange(imageSizeY, dtype=np.float64)[:, None] - centerY
    xs = np.arange(imageSizeX, dtype=np.float64)[None, :] - centerX
    return ys * ys + xs * xs <= radius * radius


def circular(centerX, centerY, imageSizeX, imageSizeY, radius, antialiased=False):
    """Boolean disk (masks.py:108-127); ``antialiased`` uses one radial bin."""
    if antialiased:
        return radial_bins(centerX, centerY, imageSizeX, imageSizeY, radius, n_bins=1,
                           use_sparse=False)[0]
    return _disk(centerX, centerY, imageSizeX, imageSizeY, radius)


def ring(centerX, centerY, imageSizeX, imageSizeY, radius, radius_inner, antialiased=False):
    """Boolean annulus: inside ``radius`` and not inside ``radius_inner`` (masks.py:130-157)."""
    if antialiased:
        return radial_bins(centerX, centerY, imageSizeX, imageSizeY, radius=radius,
                           radius_inner=radius_inner, n_bins=1, use_sparse=False)[0]
    return (_disk(centerX, centerY, imageSizeX, imageSizeY, radius)
            & ~_disk(centerX, centerY, imageSizeX, imageSizeY, radius_inner))


def gradient_x(imageSizeX, imageSizeY, dtype=np.float32):
    """Column index of every pixel (masks.py:415-418)."""
    return np.broadcast_to(np.arange(imageSizeX, dtype=dtype)[None, :],
                           (imageSizeY, imageSizeX)).copy()


def gradient_y(imageSizeX, imageSizeY, dtype=np.float32):
    """Row index of every pixel (masks.py:421-422)."""
    return np.broadcast_to(np.arange(imageSizeY, dtype=dtype)[:, None],
                           (imageSizeY, imageSizeX)).copy()


def polar_map(centerX, centerY, imageSizeX, imageSizeY, stretchY=1., angle=0.):
    """(radius, angle) maps, angle = arctan2(dy, dx) (masks.py:222-263)."""
    dy = (np.arange(imageSizeY)[:, None] - centerY) * np.ones((1, imageSizeX))
    dx = (np.arange(imageSizeX)[None, :] - centerX) * np.ones((imageSizeY, 1))
    if stretchY != 1.0 or angle != 0.:
        dy, dx = ((dy * np.cos(angle) - dx * np.sin(angle)) / stretchY,
                  dx * np.cos(angle) + dy * np.sin(angle))
    return np.sqrt(dy ** 2 + dx ** 2), np.arctan2(dy, dx)


def bounding_radius(centerX, centerY, imageSizeX, imageSizeY):
    """Radius about the centre that covers the whole frame (masks.py:281-287)."""
    dy = max(centerY, imageSizeY - centerY)
    dx = max(centerX, imageSizeX - centerX)
    return int(np.ceil(np.sqrt(dy ** 2 + dx ** 2))) + 1


def radial_bins(centerX, centerY, imageSizeX, imageSizeY, radius=None, radius_inner=0,
                n_bins=None, normalize=False, use_sparse=None, dtype=None):
    """Antialiased, overlapping rings that sum to one (masks.py:290-353).

    ``use_sparse`` True returns a list-like stack of scipy CSR rows reshaped on demand by
    MaskContainer; here the dense stack is returned for False/None-dense, and a
    ``scipy.sparse.csr_matrix`` of shape (n_bins, sy*sx) wrapped in SparseStack for sparse.
    """
    if radius is None:
        radius = bounding_radius(centerX, centerY, imageSizeX, imageSizeY)
    if n_bins is None:
        n_bins = int(np.round(radius - radius_inner))
    r, _ = polar_map(centerX, centerY, imageSizeX, imageSizeY)
    r = r.reshape(-1)
    width = (radius - radius_inner) / n_bins
    bin_area = np.pi * (radius ** 2 - (radius - width) ** 2)
    if use_sparse is None:
        use_sparse = bin_area / (imageSizeX * imageSizeY) < 0.1
    centres = np.linspace(radius_inner, radius - width, n_bins) + width / 2
    rows = []
    for r0 in centres:
        vals = np.maximum(0, np.minimum(1, width / 2 + 0.5 - np.abs(r - r0)))
        if normalize:
            total = vals.sum()
            if not np.isclose(total, 0):
                vals = vals / total
        rows.append(vals.astype(dtype))
    stack = np.stack(rows).reshape((n_bins, imageSizeY, imageSizeX))
    if radius_inner < 0.5:
        yy, xx = int(np.round(centerY)), int(np.round(centerX))
        if 0 <= yy < imageSizeY and 0 <= xx < imageSizeX:
            stack[0, yy, xx] = 1 - radius_inner
    if use_sparse:
        return SparseStack.from_dense(stack)
    return stack


class SparseStack:
    """A stack of sparse masks ``(M, *sig)`` held as scipy CSR ``(M, sig_size)``.

    Stands in for the pydata ``sparse.COO`` stacks the reference's generators return
    (that package is not a dependency here); MaskContainer treats it as 'sparse'."""

    def __init__(self, csr, sig_shape):
        import scipy.sparse as sp
        self.csr = sp.csr_matrix(csr)
        self.sig_shape = tuple(sig_shape)

    @classmethod
    def from_dense(cls, stack):
        import scipy.sparse as sp
        stack = np.asarray(stack)
        return cls(sp.csr_matrix(stack.reshape((stack.shape[0], -1))), stack.shape[1:])

    @property
    def shape(self):
        return (self.csr.shape[0],) + self.sig_shape

    @property
    def dtype(self):
        return self.csr.dtype

    def __len__(self):
        return self.csr.shape[0]

    def todense(self):
        return np.asarray(self.csr.todense()).reshape(self.shape)


def sparse_template_multi_stack(mask_index, offsetX, offsetY, template, imageSizeX, imageSizeY):
    """Stamp ``template`` into mask ``mask_index[i]`` at (offsetY[i], offsetX[i]), clipped
    (masks.py:55-83)."""
    import scipy.sparse as sp
    fy, fx = template.shape
    n = int(max(mask_index) + 1)
    rows, cols, vals = [], [], []
    ty, tx = np.mgrid[0:fy, 0:fx]
    for mi, ox, oy in zip(mask_index, offsetX, offsetY):
        y = ty.reshape(-1) + int(oy)
        x = tx.reshape(-1) + int(ox)
        ok = (y >= 0) & (y < imageSizeY) & (x >= 0) & (x < imageSizeX)
        rows.append(np.full(int(ok.sum()), mi))
        cols.append(y[ok] * imageSizeX + x[ok])
        vals.append(template.reshape(-1)[ok])
    csr = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                        shape=(n, imageSizeY * imageSizeX))
    return SparseStack(csr, (imageSizeY, imageSizeX))


def sparse_circular_multi_stack(mask_index, centerX, centerY, imageSizeX, imageSizeY, radius):
    """One small disk per entry, as a sparse stack (masks.py:86-105)."""
    bbox = int(2 * np.ceil(radius) + 1)
    c = int((bbox - 1) // 2)
    template = circular(centerX=c, centerY=c, imageSizeX=bbox, imageSizeY=bbox, radius=radius)
    return sparse_template_multi_stack(
        mask_index=mask_index,
        offsetX=np.array(centerX, dtype=int) - c, offsetY=np.array(centerY, dtype=int) - c,
        template=template, imageSizeX=imageSizeX, imageSizeY=imageSizeY)


def rectangular(X, Y, Width, Height, imageSizeX, imageSizeY):
    """Boolean rectangle with corner (X, Y); negative extents flip it (masks.py:370-411)."""
    m = np.zeros((imageSizeY, imageSizeX), dtype=bool)
    if Height * Width > 0:
        y0, y1 = min(Y, Y + Height), max(Y, Y + Height)
        x0, x1 = min(X, X + Width), max(X, X + Width)
    elif Height > 0 and Width < 0:
        y0, y1, x0, x1 = Y, Y + Height, X + Width, X
    elif Height < 0 and Width > 0:
        y0, y1, x0, x1 = Y + Height, Y, X, X + Width
    else:
        return m
    y0, y1, x0, x1 = int(y0), int(y1), int(x0), int(x1)
    m[max(0, y0):min(y1 + 1, imageSizeY), max(0, x0):min(x1 + 1, imageSizeX)] = True
    return m


def is_sparse(a):
    import scipy.sparse as sp
    return isinstance(a, SparseStack) or sp.issparse(a) or (
        type(a).__module__.split('.')[0] == 'sparse' and hasattr(a, 'todense'))


def to_dense(a):
    if isinstance(a, SparseStack):
        return a.todense()
    if hasattr(a, 'toarray'):
        return np.asarray(a.toarray())
    if hasattr(a, 'todense'):
        return np.asarray(a.todense())
    return np.asarray(a)
