"""Mask generators, restated from the reference ``src/libertem/masks.py``.

Test infrastructure (see oracle/__init__.py).  Pinned by tests/golden/masks_gen.npz.
"""
import numpy as np


def circular(centerX, centerY, imageSizeX, imageSizeY, radius):
    """bool disk ``x^2+y^2 <= r^2`` -- masks.py:18-52 (_make_circular_mask, non-antialiased)."""
    yy, xx = np.ogrid[-centerY:imageSizeY - centerY, -centerX:imageSizeX - centerX]
    return yy * yy + xx * xx <= radius * radius


def ring(centerX, centerY, imageSizeX, imageSizeY, radius, radius_inner):
    """bool ring = outer & ~inner -- masks.py:130-157."""
    outer = circular(centerX, centerY, imageSizeX, imageSizeY, radius)
    inner = circular(centerX, centerY, imageSizeX, imageSizeY, radius_inner)
    return outer & ~inner


def gradient_x(imageSizeX, imageSizeY, dtype=np.float32):
    """column index as float -- masks.py:415-418."""
    return np.tile(np.arange(imageSizeX).astype(dtype), imageSizeY).reshape(imageSizeY, imageSizeX)


def gradient_y(imageSizeX, imageSizeY, dtype=np.float32):
    """row index as float -- masks.py:421-422."""
    return gradient_x(imageSizeY, imageSizeX, dtype).transpose()


def make_polar(cartesians):
    """(y, x) -> (r, phi) with phi = arctan2(y, x) -- utils/__init__.py:27-44."""
    result = np.zeros_like(cartesians, dtype=np.float64)
    yy = cartesians[..., 0]
    xx = cartesians[..., 1]
    result[..., 0] = np.sqrt(yy ** 2 + xx ** 2)
    result[..., 1] = np.arctan2(yy, xx)
    return result


def polar_map(centerX, centerY, imageSizeX, imageSizeY, stretchY=1., angle=0.):
    """radius / angle map -- masks.py:222-263."""
    y, x = np.mgrid[0:imageSizeY, 0:imageSizeX]
    dy = y - centerY
    dx = x - centerX
    if stretchY != 1.0 or angle != 0.:
        (dy, dx) = (
            (dy * np.cos(angle) - dx * np.sin(angle)) / stretchY,
            dx * np.cos(angle) + dy * np.sin(angle),
        )
    polars = make_polar(np.stack((dy.flatten(), dx.flatten())).T)
    return (polars[:, 0].reshape((imageSizeY, imageSizeX)),
            polars[:, 1].reshape((imageSizeY, imageSizeX)))


def bounding_radius(centerX, centerY, imageSizeX, imageSizeY):
    """masks.py:281-287."""
    dy = max(centerY, imageSizeY - centerY)
    dx = max(centerX, imageSizeX - centerX)
    return int(np.ceil(np.sqrt(dy ** 2 + dx ** 2))) + 1


def radial_bins(centerX, centerY, imageSizeX, imageSizeY, radius=None, radius_inner=0,
                n_bins=None, normalize=False, dtype=None):
    """dense stack of antialiased overlapping rings -- masks.py:290-353 (dense branch;
    the sparse branch holds the same values at the non-zero positions)."""
    if radius is None:
        radius = bounding_radius(centerX, centerY, imageSizeX, imageSizeY)
    if n_bins is None:
        n_bins = int(np.round(radius - radius_inner))
    r, _ = polar_map(centerX, centerY, imageSizeX, imageSizeY)
    r = r.flatten()
    width = (radius - radius_inner) / n_bins
    slices = []
    for r0 in np.linspace(radius_inner, radius - width, n_bins) + width / 2:
        diff = np.abs(r - r0)
        vals = np.maximum(0, np.minimum(1, width / 2 + 0.5 - diff))
        if normalize:
            s = vals.sum()
            if not np.isclose(s, 0):
                vals /= s
        slices.append(vals.reshape((imageSizeY, imageSizeX)).astype(dtype))
    if radius_inner < 0.5:
        yy = int(np.round(centerY))
        xx = int(np.round(centerX))
        if 0 <= yy < imageSizeY and 0 <= xx < imageSizeX:
            slices[0][yy, xx] = 1 - radius_inner
    return np.stack(slices)
