"""MaskContainer: lazy mask factories -> stacked masks -> per-sig-slice, cached, conditioned
for the dot product.  Mirrors the interface of the reference
``libertem.common.container.MaskContainer`` (src/libertem/common/container.py:97-339):
same constructor arguments, ``use_sparse`` resolution, ``computed_masks``, ``dtype``,
``len()``, ``get`` / ``get_for_sig_slice`` / ``get_for_idx``, and caches that are dropped on
pickling.  On top of that it serves *device* forms for the CUDA kernels:

* ``get_device_dense(sig_slice)``  -> torch float32/float64 ``(M, K_tile)`` row-major, i.e. the
  physical layout of the reference's F-ordered ``(K_tile, M)`` matrix (container.py:86-91)
* ``get_device_csc(sig_slice)``    -> (indptr, indices, values) int32/int32/float32 tensors:
  per mask the ascending pixel list (the scipy CSC of the ``(K_tile, M)`` matrix)
"""
import logging

import numpy as np

from . import Slice, Shape
from ..masks import SparseStack, is_sparse, to_dense

log = logging.getLogger(__name__)

_SPARSE_NAMES = ('sparse.pydata', 'sparse.pydata.GCXS', 'scipy.sparse', 'scipy.sparse.csc',
                 'scipy.sparse.csr')


class MaskContainer:
    def __init__(self, mask_factories, dtype=None, use_sparse=None, count=None, backend=None,
                 default_sparse='scipy.sparse'):
        self.mask_factories = mask_factories
        self._length = count
        self._dtype = dtype
        self._computed_masks = None
        self.backend = 'numpy' if backend is None else backend
        self._slice_cache = {}
        self._device_cache = {}
        self._default_sparse = default_sparse
        # container.py:150-177: resolve the sparse mode as far as possible up front
        if use_sparse is True:
            self._use_sparse = default_sparse
        elif use_sparse is False:
            self._use_sparse = False
        elif isinstance(use_sparse, str) and use_sparse.lower().startswith(
                ('scipy.sparse', 'sparse.pydata')):
            self._use_sparse = use_sparse
        elif use_sparse is None:
            self._use_sparse = None     # decided when the masks exist
        else:
            raise ValueError(f'use_sparse not an allowed value: {use_sparse}')
        self.validate_mask_functions()

    # -- pickling: never ship computed masks / caches (container.py:181-185) --------------
    def __getstate__(self):
        state = dict(self.__dict__)
        state['_slice_cache'] = {}
        state['_device_cache'] = {}
        state['_computed_masks'] = None
        return state

    def validate_mask_functions(self):
        import cloudpickle
        fns = self.mask_factories
        if callable(fns):
            fns = [fns]
        for fn in fns:
            size = len(cloudpickle.dumps(fn))
            if size > 2 ** 20:
                log.warning('Mask factory size %s larger than warning limit %s, may be '
                            'inefficient' % (size, 2 ** 20))

    def __len__(self):
        if self._length is not None:
            return self._length
        if not callable(self.mask_factories):
            return len(self.mask_factories)
        return len(self.computed_masks)

    @property
    def dtype(self):
        if self._dtype is None:
            return self.computed_masks.dtype
        return self._dtype

    @property
    def use_sparse(self):
        if self._use_sparse is None:
            self._use_sparse = (self._default_sparse if is_sparse(self.computed_masks)
                                else False)
        return self._use_sparse

    @property
    def computed_masks(self):
        if self._computed_masks is None:
            self._computed_masks = self._compute_masks()
        return self._computed_masks

    def _compute_masks(self):
        """Call the factories and stack (container.py:260-314).  Returns a dense
        ``(M, *sig)`` ndarray or, in sparse mode, a SparseStack."""
        import scipy.sparse as sp
        pieces = []
        if callable(self.mask_factories):
            pieces.append(self.mask_factories())
        else:
            for f in self.mask_factories:
                m = f()
                if sp.issparse(m):
                    m = SparseStack(sp.csr_matrix(m.reshape((1, -1))), m.shape)
                elif not is_sparse(m):
                    m = np.asarray(m)
                    m = m.reshape((1,) + m.shape)
                elif not isinstance(m, SparseStack):
                    d = to_dense(m)
                    m = SparseStack.from_dense(d.reshape((1,) + d.shape))
                pieces.append(m)
        normalised = []
        for m in pieces:
            if is_sparse(m) and not isinstance(m, SparseStack):
                m = SparseStack.from_dense(to_dense(m))
            normalised.append(m)
        all_sparse = all(isinstance(m, SparseStack) for m in normalised)
        use_sparse = self._use_sparse
        if use_sparse is None:
            use_sparse = self._default_sparse if all_sparse else False
        if use_sparse is not False:
            import scipy.sparse as sp
            stacks = [m if isinstance(m, SparseStack) else SparseStack.from_dense(np.asarray(m))
                      for m in normalised]
            dt = np.result_type(*[s.dtype for s in stacks])
            return SparseStack(sp.vstack([s.csr.astype(dt) for s in stacks]).tocsr(),
                               stacks[0].sig_shape)
        return np.concatenate([to_dense(m) for m in normalised])

    # -- reference-compatible host views ----------------------------------------------------
    def get_for_idx(self, scheme, idx, *args, **kwargs):
        return self._get(scheme[idx], *args, **kwargs)

    def get_for_sig_slice(self, sig_slice, *args, **kwargs):
        return self._get(sig_slice, *args, **kwargs)

    def get(self, key, dtype=None, sparse_backend=None, transpose=True, backend=None):
        if not isinstance(key, Slice):
            raise TypeError('MaskContainer.get() can only be called with '
                            'DataTile/Slice/Partition instances')
        return self._get(key.discard_nav(), dtype, sparse_backend, transpose, backend)

    def _get(self, slice_, dtype=None, sparse_backend=None, transpose=True, backend=None):
        return self.get_masks_for_slice(slice_, dtype=dtype, sparse_backend=sparse_backend,
                                        transpose=transpose, backend=backend)

    def _dense_stack_for(self, slice_):
        """dense ``(M, K_tile)`` view of the stack restricted to the sig slice"""
        stack = self.computed_masks
        dense = stack.todense() if isinstance(stack, SparseStack) else stack
        m = slice_.get(dense, sig_only=True)
        return m.reshape((dense.shape[0], -1))

    def get_masks_for_slice(self, slice_, dtype=None, sparse_backend=None, transpose=True,
                            backend=None):
        """Host-side conditioned masks like the reference's slicer (container.py:74-94):
        dense -> ndarray ``(K, M)`` (transpose=True, F-ordered) cast to dtype;
        scipy.sparse[.csr|.csc] -> scipy matrix of the same orientation."""
        import scipy.sparse as sp
        if dtype is None:
            dtype = self.dtype
        if sparse_backend is None:
            sparse_backend = self.use_sparse
        key = (np.dtype(dtype).str, sparse_backend, transpose, slice_)
        hit = self._slice_cache.get(key)
        if hit is not None:
            return hit
        m = self._dense_stack_for(slice_)
        if transpose:
            m = m.T
        if sparse_backend is False:
            res = m.astype(dtype)
        elif sparse_backend == 'scipy.sparse.csc':
            res = sp.csc_matrix(m.astype(dtype))
        elif sparse_backend in ('scipy.sparse', 'scipy.sparse.csr'):
            res = sp.csr_matrix(m.astype(dtype))
        elif sparse_backend.startswith('sparse.pydata'):
            # no pydata-sparse dependency here: serve the equivalent scipy COO
            res = sp.coo_matrix(m.astype(dtype))
        else:
            raise ValueError(f'sparse_backend {sparse_backend} not implemented')
        self._slice_cache[key] = res
        return res

    # -- device forms for the CUDA kernels ----------------------------------------------------
    def get_device_dense(self, slice_, device, dtype=np.float32):
        import torch
        key = ('dense', np.dtype(dtype).str, slice_, str(device))
        hit = self._device_cache.get(key)
        if hit is None:
            m = np.ascontiguousarray(self._dense_stack_for(slice_).astype(dtype))
            if m.dtype.kind == 'c':
                # complex masks on real data = interleaved (re, im) real rows: the (F, 2M)
                # float result *is* the complex (F, M) result in memory
                fl = np.float32 if m.dtype == np.complex64 else np.float64
                m = np.ascontiguousarray(
                    np.stack([m.real, m.imag], axis=1).reshape((2 * m.shape[0], -1)).astype(fl))
            hit = torch.from_numpy(m).to(device)
            self._device_cache[key] = hit
        return hit

    def get_device_dense_for_complex(self, slice_, device, compute=np.float32):
        """mask rows for COMPLEX frames read as their interleaved (re, im) float view (F, 2K):
        for mask m = c + i d the rows ``[c_k at 2k, -d_k at 2k+1]`` (real part of the product
        sum) and ``[d_k at 2k, c_k at 2k+1]`` (imaginary part), so that the (F, 2M) float result
        of the dense kernel *is* the complex (F, M) result in memory."""
        import torch
        fl = np.dtype(compute)
        key = ('dense_cplx', fl.str, slice_, str(device))
        hit = self._device_cache.get(key)
        if hit is None:
            m = np.asarray(self._dense_stack_for(slice_))
            c = np.ascontiguousarray(m.real, dtype=fl)
            d = np.ascontiguousarray(m.imag, dtype=fl) if m.dtype.kind == 'c' else np.zeros_like(c)
            M, K = c.shape
            rows = np.empty((M, 2, K, 2), dtype=fl)
            rows[:, 0, :, 0] = c
            rows[:, 0, :, 1] = -d
            rows[:, 1, :, 0] = d
            rows[:, 1, :, 1] = c
            hit = torch.from_numpy(rows.reshape(2 * M, 2 * K)).to(device)
            self._device_cache[key] = hit
        return hit

    def get_group_plan(self, slice_, device):
        """group-sparse plan for the K4 kernel (libertem_b200/group_masks.py) or None when
        the stack has no uniform group structure / is not sparse enough to pay off"""
        from .. import group_masks as gm
        key = ('group', slice_, str(device))
        if key in self._device_cache:
            return self._device_cache[key]
        stack = np.asarray(self._dense_stack_for(slice_))
        size = gm.find_groups(stack)
        plan = None
        if size is not None and size >= 2:
            support = np.any(stack.reshape(stack.shape[0] // size, size, -1) != 0, axis=1)
            # worth it when the gathered entries are few compared with columns x pixels
            if np.count_nonzero(support) * size < 0.5 * stack.shape[0] * stack.shape[1]:
                # (the 2D signal shape lets the builder look for mirror symmetry, full frames only)
                try:
                    sig_shape = tuple(int(v) for v in slice_.shape)
                    if any(int(o) != 0 for o in slice_.origin):
                        sig_shape = None
                except (AttributeError, TypeError):
                    sig_shape = None
                plan = gm.build_plan(stack, size, device, sig_shape=sig_shape)
        self._device_cache[key] = plan
        return plan

    def get_device_csc(self, slice_, device):
        import torch
        import scipy.sparse as sp
        key = ('csc', slice_, str(device))
        hit = self._device_cache.get(key)
        if hit is None:
            m = self._dense_stack_for(slice_)
            csc = sp.csc_matrix(m.T.astype(np.float32))     # (K_tile, M)
            csc.sort_indices()
            hit = (torch.from_numpy(csc.indptr.astype(np.int32)).to(device),
                   torch.from_numpy(csc.indices.astype(np.int32)).to(device),
                   torch.from_numpy(csc.data.astype(np.float32)).to(device))
            self._device_cache[key] = hit
        return hit


def full_sig_slice(sig_shape):
    sig_shape = tuple(sig_shape)
    return Slice(origin=(0,) * len(sig_shape), shape=Shape(sig_shape, sig_dims=len(sig_shape)))
