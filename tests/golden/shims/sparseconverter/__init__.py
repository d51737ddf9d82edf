"""Minimal stand-in for the third-party ``sparseconverter`` package.

NOT reference code and NOT product code: it only exists so that the unmodified
reference under /root/reference/src can be imported in the build container
(which lacks sparseconverter / sparse / matplotlib and has no network) to
generate golden vectors (tests/golden/make_golden.py).  Only the names and the
behaviour the reference's dense + scipy.sparse mask paths touch are provided.
"""
import numpy as np
import scipy.sparse as sp

NUMPY = 'numpy'
NUMPY_MATRIX = 'numpy.matrix'
CUDA = 'cuda'
CUPY = 'cupy'
SPARSE_COO = 'sparse.COO'
SPARSE_GCXS = 'sparse.GCXS'
SPARSE_DOK = 'sparse.DOK'
SCIPY_COO = 'scipy.sparse.coo_matrix'
SCIPY_CSR = 'scipy.sparse.csr_matrix'
SCIPY_CSC = 'scipy.sparse.csc_matrix'
SCIPY_COO_ARRAY = 'scipy.sparse.coo_array'
SCIPY_CSR_ARRAY = 'scipy.sparse.csr_array'
SCIPY_CSC_ARRAY = 'scipy.sparse.csc_array'
CUPY_SCIPY_COO = 'cupyx.scipy.sparse.coo_matrix'
CUPY_SCIPY_CSR = 'cupyx.scipy.sparse.csr_matrix'
CUPY_SCIPY_CSC = 'cupyx.scipy.sparse.csc_matrix'

ArrayBackend = str
ArrayT = object

CPU_BACKENDS = frozenset((
    NUMPY, NUMPY_MATRIX, SPARSE_COO, SPARSE_GCXS, SPARSE_DOK, SCIPY_COO, SCIPY_CSR,
    SCIPY_CSC, SCIPY_COO_ARRAY, SCIPY_CSR_ARRAY, SCIPY_CSC_ARRAY,
))
CUPY_BACKENDS = frozenset((CUPY, CUPY_SCIPY_COO, CUPY_SCIPY_CSR, CUPY_SCIPY_CSC))
CUDA_BACKENDS = CUPY_BACKENDS | {CUDA}
BACKENDS = CPU_BACKENDS | CUDA_BACKENDS
ND_BACKENDS = frozenset((NUMPY, CUDA, CUPY, SPARSE_COO, SPARSE_GCXS, SPARSE_DOK))
D2_BACKENDS = BACKENDS - ND_BACKENDS
DENSE_BACKENDS = frozenset((NUMPY, NUMPY_MATRIX, CUDA, CUPY))
SPARSE_BACKENDS = BACKENDS - DENSE_BACKENDS


def get_backend(arr):
    import sparse
    if isinstance(arr, np.matrix):
        return NUMPY_MATRIX
    if isinstance(arr, np.ndarray):
        return NUMPY
    if isinstance(arr, sparse.COO):
        return SPARSE_COO
    if sp.issparse(arr):
        return {'coo': SCIPY_COO, 'csr': SCIPY_CSR, 'csc': SCIPY_CSC}.get(arr.format)
    return None


def get_device_class(backend):
    if backend is None:
        return 'cpu'
    return 'cuda' if backend in CUDA_BACKENDS else 'cpu'


def for_backend(arr, backend, strict=True):
    import sparse
    src = get_backend(arr)
    if backend in (NUMPY, CUDA):
        if src in (NUMPY, None):
            return np.asarray(arr) if src is None else arr
        if src == SPARSE_COO:
            return arr.todense()
        return np.asarray(arr.todense())
    if backend == SPARSE_COO:
        if src == SPARSE_COO:
            return arr
        if src == NUMPY:
            return sparse.COO.from_numpy(arr)
        return sparse.COO.from_scipy_sparse(arr)
    if backend in (SCIPY_CSR, SCIPY_CSC, SCIPY_COO):
        ctor = {SCIPY_CSR: sp.csr_matrix, SCIPY_CSC: sp.csc_matrix, SCIPY_COO: sp.coo_matrix}
        if src == SPARSE_COO:
            arr = arr.todense()
        a = np.asarray(arr) if not sp.issparse(arr) else arr
        if not sp.issparse(a) and a.ndim != 2:
            a = a.reshape((a.shape[0], -1))
        return ctor[backend](a)
    raise NotImplementedError(f'shim: for_backend to {backend}')


def make_like(arr, target, strict=False):
    return arr


def check_shape(arr, shape):
    return tuple(arr.shape) == tuple(shape)


def result_type(*args):
    items = []
    for a in args:
        if isinstance(a, str) and a in BACKENDS:
            continue
        items.append(a)
    return np.result_type(*items)


def conversion_cost(source, target):
    return 0. if source == target else 1.


def cheapest_pair(source_backends, target_backends):
    best = None
    for s in source_backends:
        for t in target_backends:
            c = conversion_cost(s, t)
            # prefer numpy->numpy
            key = (c, s != NUMPY, t != NUMPY)
            if best is None or key < best[0]:
                best = (key, (s, t))
    return best[1]
