"""Nav-space post-processing on the device (K9, libertem_b200/csrc/k9_nav.cu): what
``CoMUDF.get_results`` and ``guess_corrections`` do in numpy in the reference
(src/libertem/udf/com.py:100-142,145-295,600-717), for moments that already live in HBM."""
import ctypes

import numpy as np
import torch

from ._lib import get_lib, check


def _ptr(t):
    return None if t is None else t.data_ptr()


def com_postprocess(raw, nav_shape, cy, cx, transform, regression_mode, regression=None,
                    row_of_nav=None, valid=None):
    """raw: CUDA float32 ``(rows, >= 3)`` moments [m00, m10, m01] per (roi-compressed) scan
    position; transform: 2x2 float64 (rotation @ flip); regression_mode: -1 none, 0 mean,
    1 least-squares plane, 2 given ``regression`` (3, 2).  Returns a dict of CUDA tensors over
    the FULL scan grid -- raw_shifts / raw_com ``(n, 2)`` float32, field ``(n, 2)`` and field_y /
    field_x / magnitude / divergence / curl ``(n,)`` float64 (the reference's dtypes), NaN
    outside ``row_of_nav`` -- and 'regression' ``(3, 2)`` float64."""
    lib = get_lib()
    ny, nx = (int(v) for v in nav_shape)
    n = ny * nx
    dev = raw.device
    if raw.dtype != torch.float32 or raw.dim() != 2 or raw.shape[1] < 3 or raw.stride(1) != 1:
        raise ValueError('raw must be a float32 (rows, 3) CUDA tensor')
    f32 = dict(dtype=torch.float32, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    out = {k: torch.empty((n, 2), **f32) for k in ('raw_shifts', 'raw_com')}
    out['field'] = torch.empty((n, 2), **f64)
    out.update({k: torch.empty((n,), **f64)
                for k in ('field_y', 'field_x', 'magnitude', 'divergence', 'curl')})
    reg = torch.zeros((3, 2), dtype=torch.float64, device=dev)
    if regression_mode == 2:
        reg.copy_(torch.as_tensor(np.asarray(regression, dtype=np.float64).reshape(3, 2)))
    t = (ctypes.c_double * 4)(*[float(v) for v in np.asarray(transform, dtype=np.float64).reshape(4)])
    need = lib.ltb200_com_workspace(ny, nx)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    rows = None
    if row_of_nav is not None:
        rows = torch.as_tensor(np.ascontiguousarray(row_of_nav, dtype=np.int32)).to(dev)
    vmask = None
    if valid is not None:
        vmask = torch.as_tensor(np.ascontiguousarray(valid, dtype=np.uint8).reshape(-1)).to(dev)
    with torch.cuda.device(dev):
        check(lib.ltb200_com_postprocess(
            raw.data_ptr(), raw.stride(0), _ptr(rows), _ptr(vmask), ny, nx, float(cy), float(cx),
            t, int(regression_mode), reg.data_ptr(), out['raw_shifts'].data_ptr(),
            out['raw_com'].data_ptr(), out['field'].data_ptr(), out['field_y'].data_ptr(),
            out['field_x'].data_ptr(), out['magnitude'].data_ptr(), out['divergence'].data_ptr(),
            out['curl'].data_ptr(), ws.data_ptr(), ws.numel(),
            torch.cuda.current_stream(dev).cuda_stream))
    out['regression'] = reg
    return out


def _rotate_deg(degrees):
    rad = np.pi * degrees / 180
    return np.array([(np.cos(rad), np.sin(rad)), (-np.sin(rad), np.cos(rad))])


def guess_corrections(y_centers, x_centers, roi=None, device=None):
    """``guess_corrections`` (com.py:207-295) on the device.  The reference evaluates the RMS
    curl of the corrected field for 360 rotations x 2 flips, i.e. 720 passes over the scan grid;
    the curl of a linearly transformed field is linear in the four gradient fields of (y, x), so
    ONE pass that accumulates their 4x4 Gram matrix is enough and the 720 candidates are
    evaluated in closed form (float64).  roi: None (the reference's default window
    ``[:-1, :-1]``) or a pair of slices.  Returns ``(scan_rotation, flip_y, cy, cx)``."""
    from .udf.com import GuessResult
    lib = get_lib()
    if device is None:
        device = y_centers.device if isinstance(y_centers, torch.Tensor) and y_centers.is_cuda \
            else torch.device('cuda', torch.cuda.current_device())
    y = torch.as_tensor(y_centers).to(device=device, dtype=torch.float32).contiguous()
    x = torch.as_tensor(x_centers).to(device=device, dtype=torch.float32).contiguous()
    ny, nx = y.shape
    if roi is None:
        roi = (slice(0, -1), slice(0, -1))
    r0, r1, rs = roi[0].indices(ny)
    c0, c1, cs = roi[1].indices(nx)
    if rs != 1 or cs != 1:
        raise ValueError('guess_corrections on the device takes a window of unit-step slices')
    ws = torch.empty(296 * 17 * 8, dtype=torch.uint8, device=device)
    sums = torch.empty(17, dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    with torch.cuda.device(device):
        check(lib.ltb200_com_gradient_gram(y.data_ptr(), x.data_ptr(), ny, nx, r0, r1, c0, c1,
                                           sums.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    s = sums.cpu().numpy()
    gram = np.zeros((4, 4))
    k = 0
    for a in range(4):
        for b in range(a, 4):
            gram[a, b] = gram[b, a] = s[k]
            k += 1
    count = s[16]
    flip_m = np.array([(-1, 0), (0, 1)])
    straight = np.zeros(360)
    flipped = np.zeros(360)
    for angle in range(360):
        for flip in (True, False):
            t = _rotate_deg(angle) @ (flip_m if flip else np.eye(2))
            # curl(T f) = t00 dy/d1 + t01 dx/d1 - t10 dy/d0 - t11 dx/d0 over g = (dy/d0, dy/d1,
            # dx/d0, dx/d1)
            a = np.array([-t[1, 0], t[0, 0], -t[1, 1], t[0, 1]])
            rms = np.sqrt(max(a @ gram @ a, 0.0) / count)
            (flipped if flip else straight)[angle] = rms
    flip = bool(np.min(flipped) < np.min(straight))
    angle = int(np.argmin(flipped) if flip else np.argmin(straight))
    t = _rotate_deg(angle) @ (flip_m if flip else np.eye(2))
    tc = (ctypes.c_double * 4)(*[float(v) for v in t.reshape(4)])
    mm = torch.empty(296 * 2, dtype=torch.float64, device=device)
    hist = torch.zeros(5, dtype=torch.int64, device=device)
    nb = ctypes.c_int(0)
    with torch.cuda.device(device):
        check(lib.ltb200_com_divergence_stats(y.data_ptr(), x.data_ptr(), ny, nx, r0, r1, c0, c1,
                                              tc, 0, 0.0, mm.data_ptr(), ctypes.byref(nb), None,
                                              stream))
        m = mm[:2 * nb.value].cpu().numpy().reshape(-1, 2)
        all_range = float(max(-m[:, 0].min(), m[:, 1].max()))
        check(lib.ltb200_com_divergence_stats(y.data_ptr(), x.data_ptr(), ny, nx, r0, r1, c0, c1,
                                              tc, 1, all_range, None, ctypes.byref(nb),
                                              hist.data_ptr(), stream))
    h = hist.cpu().numpy()
    if h[0] < h[4]:
        angle += 180
    if angle > 180:
        angle -= 360
    return GuessResult(scan_rotation=int(angle), flip_y=flip, cy=float(s[14] / count),
                       cx=float(s[15] / count))
