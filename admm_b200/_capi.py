"""ctypes binding of libb200admm.so (include/b200admm.h).

The shared library is the product: it is loaded from this directory and nowhere else, and
loading fails loudly if it has not been built (python -m admm_b200.build) -- there is no
Python or CPU implementation behind these calls.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200admm.so")

F64_HOST, F32_HOST, F32_DEVICE, F64_DEVICE = 0, 1, 2, 3
COMM_ID_BYTES = 128

# symbols declared in include/b200admm.h (tests check that every one is exported)
EXPORTS = [
    "b200admm_lasso", "b200admm_enet", "b200admm_parlasso", "b200admm_free_path",
    "b200admm_lad", "b200admm_free_dense", "b200admm_bp",
    "b200admm_comm_id", "b200admm_comm_init", "b200admm_comm_destroy", "b200admm_comm_suspend", "b200admm_set_capture",
    "b200admm_last_error", "b200admm_version", "b200admm_launch_count", "b200admm_last_gram_seconds", "b200admm_last_work", "b200admm_stream", "b200admm_release_cache", "b200admm_device_info", "b200admm_set_trace",
    "b200admm_synth_f32",
    "b200admm_k_standardize_f32", "b200admm_k_gram_f32", "b200admm_k_gemv_t_f32",
    "b200admm_k_gram_plan", "b200admm_k_panel_schedule", "b200admm_k_tri_plan", "b200admm_k_lambda_grid", "b200admm_k_coarse_eig_f32", "b200admm_k_gemm_tn_f32", "b200admm_k_gemm_f64",
    "b200admm_k_chol_f32", "b200admm_k_spd_inverse_f32", "b200admm_k_fused_zu_f32",
]


class Data(C.Structure):
    _fields_ = [("n", C.c_int64), ("p", C.c_int64), ("dtype", C.c_int), ("x", C.c_void_p), ("y", C.c_void_p)]


class Opts(C.Structure):
    _fields_ = [("maxit", C.c_int), ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("rho", C.c_double)]


class Timing(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("ingest", "standardize", "gram", "eig", "factor", "iterate", "finish", "total")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Path(C.Structure):
    _fields_ = [("nlambda", C.c_int), ("nrow", C.c_int64), ("lambda_", C.POINTER(C.c_double)),
                ("niter", C.POINTER(C.c_int)), ("colptr", C.POINTER(C.c_int64)),
                ("rowidx", C.POINTER(C.c_int)), ("val", C.POINTER(C.c_double)),
                ("rho", C.c_double), ("eig", C.c_double), ("lambda0", C.c_double), ("t", Timing)]


class Dense(C.Structure):
    _fields_ = [("len", C.c_int64), ("beta", C.POINTER(C.c_double)), ("niter", C.c_int),
                ("rho", C.c_double), ("t", Timing)]


class B200AdmmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libb200admm error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m admm_b200.build` "
                              "(nvcc, sm_100a). There is no fallback implementation." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.b200admm_last_error.restype = C.c_char_p
        L.b200admm_launch_count.restype = C.c_ulonglong
        L.b200admm_last_gram_seconds.restype = C.c_double
        L.b200admm_last_work.argtypes = [C.c_void_p]
        L.b200admm_last_work.restype = None
        L.b200admm_stream.restype = C.c_void_p
        L.b200admm_release_cache.restype = None
        L.b200admm_lasso.argtypes = [C.POINTER(Data), C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                     C.POINTER(Opts), C.POINTER(Path)]
        L.b200admm_enet.argtypes = [C.POINTER(Data), C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                    C.c_double, C.POINTER(Opts), C.POINTER(Path)]
        L.b200admm_parlasso.argtypes = [C.POINTER(Data), C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                        C.c_int, C.POINTER(Opts), C.POINTER(Path)]
        L.b200admm_free_path.argtypes = [C.POINTER(Path)]
        L.b200admm_free_path.restype = None
        L.b200admm_lad.argtypes = [C.POINTER(Data), C.c_int, C.POINTER(Opts), C.POINTER(Dense)]
        L.b200admm_free_dense.argtypes = [C.POINTER(Dense)]
        L.b200admm_free_dense.restype = None
        L.b200admm_bp.argtypes = [C.POINTER(Data), C.POINTER(Opts), C.POINTER(Path)]
        L.b200admm_comm_id.argtypes = [C.c_void_p]
        L.b200admm_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b200admm_comm_destroy.restype = None
        L.b200admm_comm_suspend.argtypes = [C.c_int]
        L.b200admm_comm_suspend.restype = None
        L.b200admm_set_capture.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200admm_set_capture.restype = None
        L.b200admm_device_info.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        L.b200admm_set_trace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.b200admm_set_trace.restype = None
        L.b200admm_synth_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64,
                                         C.c_float, C.c_float, C.c_int, C.c_float]
        L.b200admm_k_standardize_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200admm_k_gram_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.b200admm_k_gram_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.b200admm_k_panel_schedule.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.b200admm_k_tri_plan.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
        L.b200admm_k_lambda_grid.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.b200admm_k_gemv_t_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.b200admm_k_coarse_eig_f32.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.c_void_p]
        L.b200admm_k_gemm_tn_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                             C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
        L.b200admm_k_gemm_f64.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float), C.c_int]
        L.b200admm_k_chol_f32.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int)]
        L.b200admm_k_spd_inverse_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]
        L.b200admm_k_fused_zu_f32.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_double, C.c_double, C.c_int, C.c_double,
                                                                 C.c_void_p, C.POINTER(C.c_float), C.c_int]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise B200AdmmError(rc, lib().b200admm_last_error().decode(errors="replace"))


def device_info():
    name = C.create_string_buffer(256)
    sms = C.c_int(0)
    mem = C.c_int64(0)
    check(lib().b200admm_device_info(name, 256, C.byref(sms), C.byref(mem)))
    return dict(name=name.value.decode(), sm_count=sms.value, mem_bytes=mem.value)


# ---------------------------------------------------------------------------------------------
# argument marshalling
# ---------------------------------------------------------------------------------------------
def _is_torch(a):
    return type(a).__module__.startswith("torch")


def make_data(x, y, want64=False):
    """Describe (x, y) for the C ABI without copying when the layout already fits.

    numpy float64 (R's layout, Fortran order) -> F64_HOST; numpy float32 -> F32_HOST;
    CUDA torch tensors (column-major: pass x as the transpose of a contiguous (p, n) tensor,
    i.e. x.t().is_contiguous()) -> F32_DEVICE / F64_DEVICE.  Returns (Data, keepalive)."""
    if _is_torch(x):
        import torch
        if not x.is_cuda:
            x = x.numpy()
            y = y.numpy() if _is_torch(y) else y
        else:
            n, p = x.shape
            if not x.t().is_contiguous():
                x = x.t().contiguous().t()
            y = y.contiguous()
            if x.dtype == torch.float32 and y.dtype == torch.float32:
                dt = F32_DEVICE
            elif x.dtype == torch.float64 and y.dtype == torch.float64:
                dt = F64_DEVICE
            else:
                raise TypeError("device x and y must both be float32 or both float64")
            torch.cuda.current_stream().synchronize()
            d = Data(n, p, dt, x.data_ptr(), y.data_ptr())
            return d, (x, y)
    x = np.asarray(x)
    y = np.asarray(y).reshape(-1)
    if x.ndim != 2:
        raise ValueError("x must be a matrix")
    if x.dtype == np.float32 and not want64:
        xa = np.asfortranarray(x)
        ya = np.ascontiguousarray(y, dtype=np.float32)
        dt = F32_HOST
    else:
        xa = np.asfortranarray(x, dtype=np.float64)
        ya = np.ascontiguousarray(y, dtype=np.float64)
        dt = F64_HOST
    n, p = xa.shape
    d = Data(n, p, dt, xa.ctypes.data, ya.ctypes.data)
    return d, (xa, ya)


def path_to_python(P):
    """b200admm_path -> (lambda, scipy CSC beta, niter, info) and free the C side."""
    from scipy.sparse import csc_matrix
    nl = P.nlambda
    lam = np.ctypeslib.as_array(P.lambda_, shape=(nl,)).copy() if nl and P.lambda_ else np.zeros(0)
    niter = np.ctypeslib.as_array(P.niter, shape=(nl,)).copy() if nl and P.niter else np.zeros(0, dtype=np.int32)
    colptr = np.ctypeslib.as_array(P.colptr, shape=(nl + 1,)).copy()
    nnz = int(colptr[-1])
    rowidx = np.ctypeslib.as_array(P.rowidx, shape=(max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(P.val, shape=(max(nnz, 1),))[:nnz].copy()
    beta = csc_matrix((val, rowidx, colptr), shape=(int(P.nrow), nl))
    info = dict(rho=P.rho, eig=P.eig, lambda0=P.lambda0, timing=P.t.as_dict())
    lib().b200admm_free_path(C.byref(P))
    return lam, beta, niter, info


class trace:
    """Context manager: record the per-iteration scalars of lambda index `which`."""

    def __init__(self, which=0, cap=20000):
        self.buf = np.zeros((cap, 5))
        self.n = C.c_int(0)
        self.which = which
        self.cap = cap

    def __enter__(self):
        lib().b200admm_set_trace(self.buf.ctypes.data, self.cap, self.which, C.byref(self.n))
        return self

    def __exit__(self, *a):
        lib().b200admm_set_trace(None, 0, -1, None)

    @property
    def rows(self):
        return self.buf[: self.n.value]


class capture:
    """Context manager: receive the standardised Gram matrix X'X (before rho is added) and X'y of the tall lasso /
    enet fits made inside the block (parity checks at sizes where X cannot go to a CPU program)."""

    def __init__(self, p):
        self.gram = np.zeros((p, p), dtype=np.float32, order="F")
        self.xy = np.zeros(p, dtype=np.float32)
        self.stats = np.zeros(2 * p + 2, dtype=np.float32)
        self.p = p

    def __enter__(self):
        lib().b200admm_set_capture(self.gram.ctypes.data, self.xy.ctypes.data, self.stats.ctypes.data)
        return self

    def __exit__(self, *a):
        lib().b200admm_set_capture(None, None, None)

    @property
    def meanX(self):
        return self.stats[:self.p]

    @property
    def scaleX(self):
        return self.stats[self.p:2 * self.p]

    @property
    def meanY(self):
        return float(self.stats[2 * self.p])

    @property
    def scaleY(self):
        return float(self.stats[2 * self.p + 1])
