"""ctypes binding of libngsb200.so (include/ngsb200.h).

The library is the product; this module only loads it and declares the prototypes.  There
is no fallback of any kind: if the shared object is missing or a call fails, an exception
is raised (the reference raises ngstd::Exception -> NgException in the same places).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libngsb200.so")

OK = 0
REAL, COMPLEX, BLOCK3 = 0, 1, 3
IP_REAL, IP_COMPLEX, IP_COMPLEX_CONJ = 0, 1, 2


class NgsbError(RuntimeError):
    """Raised for every non-zero status of the C ABI (stands in for NgException)."""


_lib = None

_vp, _sz, _i, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
_pvp = C.POINTER(C.c_void_p)

# name -> argtypes; every function returns int unless listed in _SPECIAL
PROTOTYPES = {
    "ngsb_ctx_create": [_i, _pvp],
    "ngsb_ctx_destroy": [_vp],
    "ngsb_ctx_sync": [_vp],
    "ngsb_ctx_device": [_vp, C.POINTER(_i), C.POINTER(_i)],
    "ngsb_ctx_launch_count": [_vp, C.POINTER(C.c_uint64)],
    "ngsb_ctx_set_option": [_vp, C.c_char_p, C.c_long],
    "ngsb_ctx_kernel_time": [_vp, C.c_char_p, C.POINTER(_d), C.POINTER(C.c_uint64)],
    "ngsb_ctx_kernel_time_reset": [_vp],
    "ngsb_vec_create": [_vp, _sz, _i, _pvp],
    "ngsb_vec_destroy": [_vp],
    "ngsb_vec_info": [_vp, C.POINTER(_sz), C.POINTER(_i), C.POINTER(_sz)],
    "ngsb_vec_range": [_vp, _sz, _sz, _pvp],
    "ngsb_vec_h2d": [_vp, _vp, _sz, _sz],
    "ngsb_vec_d2h": [_vp, _vp, _sz, _sz],
    "ngsb_vec_set_scalar": [_vp, C.POINTER(_d)],
    "ngsb_vec_scale": [_vp, C.POINTER(_d)],
    "ngsb_vec_set": [_vp, C.POINTER(_d), _vp],
    "ngsb_vec_axpy": [_vp, C.POINTER(_d), _vp],
    "ngsb_vec_dot": [_vp, _vp, _i, C.POINTER(_d)],
    "ngsb_vec_nrm2": [_vp, C.POINTER(_d)],
    "ngsb_scalar_create": [_vp, _pvp],
    "ngsb_scalar_destroy": [_vp],
    "ngsb_scalar_set": [_vp, C.POINTER(_d)],
    "ngsb_scalar_get": [_vp, C.POINTER(_d)],
    "ngsb_scalar_div": [_vp, _vp, _vp],
    "ngsb_scalar_neg": [_vp, _vp],
    "ngsb_scalar_copy": [_vp, _vp],
    "ngsb_vec_dot_dev": [_vp, _vp, _i, _vp],
    "ngsb_vec_axpy_dev": [_vp, _vp, _vp],
    "ngsb_vec_scale_dev": [_vp, _vp],
    "ngsb_csr_create": [_vp, _sz, _sz, _sz, _vp, _vp, _vp, _i, _pvp],
    "ngsb_csr_create_from_device": [_vp, _sz, _sz, _sz, _vp, _vp, _vp, _i, _pvp],
    "ngsb_csr_destroy": [_vp],
    "ngsb_csr_info": [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i)],
    "ngsb_csr_multadd": [_vp, C.POINTER(_d), _vp, _vp],
    "ngsb_csr_mult": [_vp, _vp, _vp],
    "ngsb_csr_reorder": [_vp, _vp, _pvp],
    "ngsb_csr_rcm": [_vp, _vp],
    "ngsb_csr_memory": [_vp, _vp, _vp, _vp],
    "ngsb_csr_archive_size": [_vp, C.POINTER(_sz)],
    "ngsb_csr_archive_write": [_vp, _vp, _sz],
    "ngsb_csr_create_from_archive": [_vp, _vp, _sz, _i, _pvp],
    "ngsb_csr_reorder_info": [_vp, _vp, _vp, _vp],
    "ngsb_csr_download": [_vp, _vp, _vp, _vp],
    "ngsb_csr_mult_bytes": [_vp, C.POINTER(_d)],
    "ngsb_csr_stream_bytes": [_vp, C.POINTER(_d), C.POINTER(C.c_uint64)],
    "ngsb_csr_layout": [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
    "ngsb_jacobi_create": [_vp, _sz, _vp, _i, _vp, _pvp],
    "ngsb_jacobi_create_from_csr": [_vp, _vp, _pvp],
    "ngsb_jacobi_destroy": [_vp],
    "ngsb_jacobi_download": [_vp, _vp],
    "ngsb_jacobi_multadd": [_vp, C.POINTER(_d), _vp, _vp],
    "ngsb_jacobi_mult": [_vp, _vp, _vp],
    "ngsb_csr_transpose": [_vp, _pvp],
    "ngsb_csr_multtransadd": [_vp, C.POINTER(_d), _vp, _vp],
    "ngsb_csr_create_symmetric": [_vp, _sz, _sz, _vp, _vp, _vp, _i, _pvp],
    "ngsb_csr_multadd_multi": [_vp, _sz, _vp, _vp, _vp],
    "ngsb_blockjacobi_create": [_vp, _sz, _vp, _vp, _pvp],
    "ngsb_blockjacobi_create_from_inverses": [_vp, _sz, _sz, _vp, _vp, _vp, _pvp],
    "ngsb_blockjacobi_destroy": [_vp],
    "ngsb_blockjacobi_info": [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)],
    "ngsb_blockjacobi_download": [_vp, _vp],
    "ngsb_blockjacobi_multadd": [_vp, _d, _vp, _vp, _i],
    "ngsb_blockjacobi_mult": [_vp, _vp, _vp, _i],
    "ngsb_cg_solve": [_vp, _vp, _vp, _vp, _d, _i, _i, _i, C.POINTER(_i), _vp, _i, C.POINTER(_i)],
    "ngsb_cg_solve_host": [_vp, _vp, _vp, _vp, _d, _i, _i, C.POINTER(_i), _vp, _i, C.POINTER(_i)],
    "ngsb_gmres_solve": [_vp, _vp, _vp, _vp, _d, _i, _i, C.POINTER(_i), _vp, _i, C.POINTER(_i)],
    "ngsb_comm_unique_id": [_vp],
    "ngsb_comm_create": [_vp, _i, _i, _vp, _pvp],
    "ngsb_comm_create_ex": [_vp, _i, _i, _vp, _vp, _vp, _i, _pvp],
    "ngsb_comm_info": [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "ngsb_comm_destroy": [_vp],
    "ngsb_parmat_create": [_vp, _vp, _vp, _vp, _pvp],
    "ngsb_parmat_create_ex": [_vp, _vp, _vp, _vp, _vp, _vp, _pvp],
    "ngsb_parmat_info": [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz), C.POINTER(_sz)],
    "ngsb_parmat_overlap_info": [_vp, C.POINTER(_i), C.POINTER(_sz), C.POINTER(_sz)],
    "ngsb_parmat_destroy": [_vp],
    "ngsb_parmat_masterdofs": [_vp, _vp],
    "ngsb_parmat_jacobi_create": [_vp, _vp, _pvp],
    "ngsb_parmat_cumulate": [_vp, _vp],
    "ngsb_parmat_mult": [_vp, _vp, _vp],
    "ngsb_parmat_dot": [_vp, _vp, _vp, _i, _i, C.POINTER(_d)],
    "ngsb_parmat_cg_solve": [_vp, _vp, _vp, _vp, _d, _i, _i, C.POINTER(_i), _vp, _i, C.POINTER(_i)],
    "ngsb_parmat_gmres_solve": [_vp, _vp, _vp, _vp, _d, _i, C.POINTER(_i), _vp, _i, C.POINTER(_i)],
}
# bootstrap all-gather callback (ngsb_allgather_fn)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)

_SPECIAL = {
    "ngsb_last_error": ([], C.c_char_p),
    "ngsb_version": ([], C.c_char_p),
    "ngsb_ctx_stream": ([_vp], _vp),
    "ngsb_vec_devptr": ([_vp], _vp),
}


def lib():
    """The loaded library; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NgsbError("libngsb200.so is not built (%s): run `make -C ngsolve_b200/csrc`; "
                            "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, args in PROTOTYPES.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int
        for name, (args, res) in _SPECIAL.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = res
        _lib = L
    return _lib


def check(status):
    if status != OK:
        raise NgsbError(lib().ngsb_last_error().decode("utf-8", "replace"))


def scal2(s):
    """python scalar -> double[2] (re, im)."""
    z = complex(s)
    return (C.c_double * 2)(z.real, z.imag)
