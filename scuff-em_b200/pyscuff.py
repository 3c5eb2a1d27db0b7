"""ctypes binding of the two in-tree shared libraries (tests / bench.py use this; it is plumbing).

  libscuff_cuda.so  — the C ABI (include/libscuff_cuda.h): CUDA kernels for sm_100a
  libscuff_b200.so  — the host-side C++ mirror of the reference API (include/libscuff_b200.h:
                      RWGGeometry::AssembleBEMMatrix[Block], HMatrix) through its extern "C" facade

There is no Python or CPU implementation of the path behind these classes: if the libraries are
missing, or no CUDA device is present, the calls raise.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB = os.path.join(_HERE, "libscuff_cuda.so")
HOST_LIB = os.path.join(_HERE, "libscuff_b200.so")


class ScuffError(RuntimeError):
    pass


class SSIArgs(C.Structure):
    """scuffcuda_ssi_args (include/libscuff_cuda.h)"""
    _fields_ = [("omega", C.c_double * 2),
                ("eps_a", C.c_double * 2), ("mu_a", C.c_double * 2), ("sign_a", C.c_double),
                ("eps_b", C.c_double * 2), ("mu_b", C.c_double * 2), ("sign_b", C.c_double),
                ("symmetric", C.c_int), ("accumulate", C.c_int),
                ("row_offset", C.c_int), ("col_offset", C.c_int),
                ("use_displacement", C.c_int), ("displacement", C.c_double * 3),
                ("force_taylor_duffy", C.c_int), ("far_only", C.c_int),
                ("col_begin", C.c_int), ("col_end", C.c_int), ("gba_a", C.c_int), ("gba_b", C.c_int), ("reserved", C.c_int * 3)]


_cuda = None
_host = None


def cuda_lib():
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB):
            raise ScuffError(f"{CUDA_LIB} not built: run python -c 'import __graft_entry__ as g; g.build()'")
        L = C.CDLL(CUDA_LIB, mode=C.RTLD_GLOBAL)
        L.scuffcuda_last_error.restype = C.c_char_p
        L.scuffcuda_last_error.argtypes = [C.c_void_p]
        L.scuffcuda_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.scuffcuda_destroy.argtypes = [C.c_void_p]
        L.scuffcuda_set_num_surfaces.argtypes = [C.c_void_p, C.c_int]
        L.scuffcuda_set_surface.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.scuffcuda_matrix_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.scuffcuda_matrix_destroy.argtypes = [C.c_void_p, C.c_int]
        L.scuffcuda_matrix_zero.argtypes = [C.c_void_p, C.c_int]
        L.scuffcuda_matrix_copy_block.argtypes = [C.c_void_p] + [C.c_int] * 7
        L.scuffcuda_matrix_symmetrize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.scuffcuda_matrix_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
        L.scuffcuda_matrix_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
        L.scuffcuda_matrix_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.scuffcuda_lu_factor.argtypes = [C.c_void_p, C.c_int]
        L.scuffcuda_lu_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.scuffcuda_assemble_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(SSIArgs), C.c_int]
        L.scuffcuda_get_ppi_counts.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.scuffcuda_get_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.scuffcuda_synchronize.argtypes = [C.c_void_p]
        L.scuffcuda_fippi_preload.argtypes = [C.c_void_p, C.c_char_p]
        L.scuffcuda_fippi_store.argtypes = [C.c_void_p, C.c_char_p]
        L.scuffcuda_fippi_num_records.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.scuffcuda_fippi_clear.argtypes = [C.c_void_p]
        L.scuffcuda_get_pair_counts.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.scuffcuda_get_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.scuffcuda_get_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.scuffcuda_host_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.scuffcuda_host_free.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffcuda_selftest_math.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.scuffcuda_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.scuffcuda_gbar_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                            C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.scuffcuda_gbar_destroy.argtypes = [C.c_void_p, C.c_int]
        L.scuffcuda_gbar_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.scuffcuda_gbar_get_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.scuffcuda_gbar_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.scuffcuda_gbar_ewald.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _cuda = L
    return _cuda


def host_lib():
    global _host
    if _host is None:
        cuda_lib()
        if not os.path.exists(HOST_LIB):
            raise ScuffError(f"{HOST_LIB} not built: run python -c 'import __graft_entry__ as g; g.build()'")
        L = C.CDLL(HOST_LIB)
        L.scuffb200_last_error.restype = C.c_char_p
        L.scuffb200_context.restype = C.c_void_p
        L.scuffb200_geometry_open.restype = C.c_void_p
        L.scuffb200_geometry_open.argtypes = [C.c_char_p]
        L.scuffb200_geometry_close.argtypes = [C.c_void_p]
        L.scuffb200_geometry_info.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffb200_surface_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for f in ("scuffb200_surface_label", "scuffb200_region_label", "scuffb200_region_material"):
            getattr(L, f).restype = C.c_char_p
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.scuffb200_surface_arrays.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.scuffb200_eps_mu.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.scuffb200_transform.argtypes = [C.c_void_p, C.c_char_p]
        L.scuffb200_untransform.argtypes = [C.c_void_p]
        L.scuffb200_zero_region.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.scuffb200_matrix_alloc.restype = C.c_void_p
        L.scuffb200_matrix_alloc.argtypes = [C.c_void_p, C.c_int]
        L.scuffb200_matrix_new.restype = C.c_void_p
        L.scuffb200_matrix_new.argtypes = [C.c_int, C.c_int, C.c_int]
        L.scuffb200_matrix_free.argtypes = [C.c_void_p]
        L.scuffb200_matrix_data.restype = C.c_void_p
        L.scuffb200_matrix_data.argtypes = [C.c_void_p]
        L.scuffb200_matrix_shape.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffb200_matrix_sync_host.argtypes = [C.c_void_p]
        L.scuffb200_matrix_set_entry.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.scuffb200_matrix_add_entry.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.scuffb200_matrix_get_entry.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.scuffb200_matrix_zero.argtypes = [C.c_void_p]
        L.scuffb200_matrix_zero_block.argtypes = [C.c_void_p] + [C.c_int] * 4
        L.scuffb200_matrix_insert_block.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6
        L.scuffb200_matrix_mark_host_dirty.argtypes = [C.c_void_p]
        L.scuffb200_matrix_state.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.scuffb200_matrix_device_handle.argtypes = [C.c_void_p]
        L.scuffb200_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int]
        L.scuffb200_assemble_columns.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
        L.scuffb200_matrix_symmetrize.argtypes = [C.c_void_p]
        L.scuffb200_get_overlap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.scuffb200_get_overlap.restype = C.c_double
        L.scuffb200_surface_zeta.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.scuffb200_assemble_inner_cells.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                                     C.POINTER(C.c_double), C.c_int, C.c_int]
        L.scuffb200_assemble_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                               C.c_void_p, C.c_int, C.c_int]
        L.scuffb200_ppi_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffb200_create_abmb.restype = C.c_void_p
        L.scuffb200_create_abmb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.scuffb200_destroy_abmb.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffb200_assemble_block_acc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.scuffb200_lu_factorize.argtypes = [C.c_void_p]
        L.scuffb200_lu_solve.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffb200_preload_cache.argtypes = [C.c_char_p]
        L.scuffb200_store_cache.argtypes = [C.c_char_p]
        L.scuffb200_set_device.argtypes = [C.c_int]
        L.scuffb200_cache_records.restype = C.c_long
        L.scuffb200_vector_new.restype = C.c_void_p
        L.scuffb200_vector_new.argtypes = [C.c_int, C.c_int]
        L.scuffb200_vector_free.argtypes = [C.c_void_p]
        L.scuffb200_planewave_new.restype = C.c_void_p
        L.scuffb200_planewave_new.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        L.scuffb200_pointsource_new.restype = C.c_void_p
        L.scuffb200_pointsource_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p]
        L.scuffb200_incfield_set_next.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffb200_incfield_free.argtypes = [C.c_void_p]
        L.scuffb200_assemble_rhs.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.scuffb200_assemble_rhs_kbloch.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.scuffb200_lu_solve_vector.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffb200_vector_copy.argtypes = [C.c_void_p, C.c_void_p]
        L.scuffb200_get_pft_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.scuffb200_surface_origin.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.scuffb200_get_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.scuffb200_get_dsipft_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_double, C.c_int, C.c_char_p, C.c_void_p]
        L.scuffb200_incfield_eval.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.scuffb200_get_dyadic_gfs.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.scuffb200_get_rf_matrix.restype = C.c_void_p
        L.scuffb200_get_rf_matrix.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
        _host = L
    return _host


def _check(rc):
    if rc:
        raise ScuffError(host_lib().scuffb200_last_error().decode())


def set_device(dev):
    _check(host_lib().scuffb200_set_device(int(dev)))


def context():
    """the process-wide scuffcuda_context* (as int) the C++ mirror uses"""
    p = host_lib().scuffb200_context()
    if not p:
        raise ScuffError(host_lib().scuffb200_last_error().decode())
    return p


class HMatrix:
    """scuff::HMatrix (column-major); .array is a numpy COPY of the host data, shape (NR, NC)"""

    def __init__(self, ptr, owner=True):
        self.ptr = ptr
        self.owner = owner
        sh = (C.c_int * 3)()
        host_lib().scuffb200_matrix_shape(ptr, sh)
        self.NR, self.NC, self.is_real = sh[0], sh[1], bool(sh[2])

    @classmethod
    def new(cls, nr, nc, is_real=False):
        p = host_lib().scuffb200_matrix_new(nr, nc, int(is_real))
        if not p:
            raise ScuffError(host_lib().scuffb200_last_error().decode())
        return cls(p)

    @property
    def array(self):
        _check(host_lib().scuffb200_matrix_sync_host(self.ptr))
        data = host_lib().scuffb200_matrix_data(self.ptr)
        n = self.NR * self.NC
        if self.is_real:
            buf = (C.c_double * n).from_address(data)
            a = np.frombuffer(buf, dtype=np.float64)
        else:
            buf = (C.c_double * (2 * n)).from_address(data)
            a = np.frombuffer(buf, dtype=np.complex128)
        return np.array(a.reshape(self.NC, self.NR).T)   # column-major storage; copy (the HMatrix may be freed)

    def SetEntry(self, r, c, v):
        _check(host_lib().scuffb200_matrix_set_entry(self.ptr, r, c, float(np.real(v)), float(np.imag(v))))

    def AddEntry(self, r, c, v):
        _check(host_lib().scuffb200_matrix_add_entry(self.ptr, r, c, float(np.real(v)), float(np.imag(v))))

    def GetEntry(self, r, c):
        out = (C.c_double * 2)()
        _check(host_lib().scuffb200_matrix_get_entry(self.ptr, r, c, out))
        return complex(out[0], out[1])

    def Zero(self):
        _check(host_lib().scuffb200_matrix_zero(self.ptr))

    def ZeroBlock(self, ro, nr, co, nc):
        _check(host_lib().scuffb200_matrix_zero_block(self.ptr, ro, nr, co, nc))

    def InsertBlock(self, B, RowOffset, ColOffset, NRB=None, NCB=None, BRowOffset=0, BColOffset=0):
        _check(host_lib().scuffb200_matrix_insert_block(self.ptr, B.ptr, RowOffset, ColOffset, B.NR if NRB is None else NRB,
                                                         B.NC if NCB is None else NCB, BRowOffset, BColOffset))

    def host_view(self):
        """writable numpy view of the HOST array (column-major: index [col, row]); call MarkHostDirty() after writing"""
        data = host_lib().scuffb200_matrix_data(self.ptr)
        n = self.NR * self.NC
        buf = (C.c_double * (n if self.is_real else 2 * n)).from_address(data)
        return np.frombuffer(buf, dtype=np.float64 if self.is_real else np.complex128).reshape(self.NC, self.NR)

    def MarkHostDirty(self):
        _check(host_lib().scuffb200_matrix_mark_host_dirty(self.ptr))

    def state(self):
        st = (C.c_int * 3)()
        host_lib().scuffb200_matrix_state(self.ptr, st)
        return dict(HostStale=bool(st[0]), DeviceStale=bool(st[1]), OnDevice=bool(st[2]))

    @property
    def device_handle(self):
        return host_lib().scuffb200_matrix_device_handle(self.ptr)

    def symmetrize(self):
        _check(host_lib().scuffb200_matrix_symmetrize(self.ptr))

    def device_tensor(self):
        """zero-copy torch view (float64, length NR*NC*(1|2), column-major) of the HBM-resident matrix"""
        import torch
        dptr = C.c_void_p()
        cuda_lib().scuffcuda_matrix_device_ptr(context(), self.device_handle, C.byref(dptr))
        n = self.NR * self.NC * (1 if self.is_real else 2)

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (dptr.value, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda")

    def lu_factorize(self):
        _check(host_lib().scuffb200_lu_factorize(self.ptr))

    def lu_solve(self, X):
        _check(host_lib().scuffb200_lu_solve(self.ptr, X.ptr))

    def lu_solve_vector(self, X):
        """HMatrix::LUSolve(HVector *): solves in HBM, the vector is not copied through the host"""
        _check(host_lib().scuffb200_lu_solve_vector(self.ptr, X.ptr))

    def __del__(self):
        if getattr(self, "owner", False) and self.ptr and _host is not None:
            _host.scuffb200_matrix_free(self.ptr)
            self.ptr = None


class HVector(HMatrix):
    """scuff::HVector (an N x 1 HMatrix with the same device twin); .array has shape (N,)"""

    @classmethod
    def new(cls, n, is_real=False):
        p = host_lib().scuffb200_vector_new(int(n), int(is_real))
        if not p:
            raise ScuffError(host_lib().scuffb200_last_error().decode())
        return cls(p)

    @property
    def array(self):
        return HMatrix.array.fget(self)[:, 0]

    def set(self, values):
        self.host_view()[0, :] = np.asarray(values)
        self.MarkHostDirty()

    def Copy(self, V):
        _check(host_lib().scuffb200_vector_copy(self.ptr, V.ptr))

    def __del__(self):
        if getattr(self, "owner", False) and self.ptr and _host is not None:
            _host.scuffb200_vector_free(self.ptr)
            self.ptr = None


class IncField:
    """scuff::IncField chain element (PlaneWave evaluated on the device; PointSource by its own GetFields at the device's points)"""

    def __init__(self, ptr):
        self.ptr = ptr
        self.next = None

    def eval(self, Omega, X, eps=1.0, mu=1.0, gradients=False):
        """GetFields (and GetFieldGradients) of this field alone in a medium (eps, mu): (NX, 6) [and (NX, 3, 6)] complex"""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1, 3))
        em = np.array([np.real(eps), np.imag(eps), np.real(mu), np.imag(mu)], dtype=np.float64)
        eh = np.zeros((X.shape[0], 6), dtype=np.complex128)
        deh = np.zeros((X.shape[0], 3, 6), dtype=np.complex128) if gradients else None
        _check(host_lib().scuffb200_incfield_eval(self.ptr, float(np.real(Omega)), float(np.imag(Omega)), em.ctypes.data, X.shape[0], X.ctypes.data,
                                                   eh.ctypes.data, deh.ctypes.data if gradients else None))
        return (eh, deh) if gradients else eh

    def chain(self, other):
        host_lib().scuffb200_incfield_set_next(self.ptr, other.ptr if other is not None else None)
        self.next = other
        return self

    def __del__(self):
        if self.ptr and _host is not None:
            _host.scuffb200_incfield_free(self.ptr)
            self.ptr = None


def PlaneWave(E0, nHat, label=None):
    e0 = np.ascontiguousarray(np.asarray(E0, dtype=np.complex128)).view(np.float64)
    n = np.ascontiguousarray(np.asarray(nHat, dtype=np.float64))
    p = host_lib().scuffb200_planewave_new(e0.ctypes.data, n.ctypes.data, (label or "").encode())
    if not p:
        raise ScuffError(host_lib().scuffb200_last_error().decode())
    return IncField(p)


def PointSource(X0, P, type=0, label=None):
    x0 = np.ascontiguousarray(np.asarray(X0, dtype=np.float64))
    pp = np.ascontiguousarray(np.asarray(P, dtype=np.complex128)).view(np.float64)
    p = host_lib().scuffb200_pointsource_new(x0.ctypes.data, pp.ctypes.data, int(type), (label or "").encode())
    if not p:
        raise ScuffError(host_lib().scuffb200_last_error().decode())
    return IncField(p)


class RWGGeometry:
    """scuff::RWGGeometry from a .scuffgeo file"""

    def AllocateRHSVector(self, PureImagFreq=False):
        return HVector.new(self.TotalBFs, bool(PureImagFreq) and self.LDim == 0)

    def AssembleRHSVector(self, Omega, IF, RHS=None, kBloch=None):
        if RHS is None:
            RHS = self.AllocateRHSVector()
        if kBloch is not None:
            kb = (C.c_double * 2)(float(kBloch[0]), float(kBloch[1]))
            _check(host_lib().scuffb200_assemble_rhs_kbloch(self.ptr, float(np.real(Omega)), float(np.imag(Omega)), kb, IF.ptr, RHS.ptr))
        else:
            _check(host_lib().scuffb200_assemble_rhs(self.ptr, float(np.real(Omega)), float(np.imag(Omega)), IF.ptr, RHS.ptr))
        return RHS

    def GetPFTMatrix(self, KN, Omega, IF=None, Interior=False, EMTPFTIMethod=0):
        out = np.zeros((self.NumSurfaces, 8))
        _check(host_lib().scuffb200_get_pft_matrix(self.ptr, KN.ptr, float(np.real(Omega)), float(np.imag(Omega)),
                                                    IF.ptr if IF is not None else None, int(Interior), int(EMTPFTIMethod), out.ctypes.data))
        return out

    def GetDSIPFTMatrix(self, KN, Omega, IF=None, DSIRadius=0.0, DSIPoints=302, DSIMesh=None):
        """GetPFTMatrix with PFTMethod = SCUFF_PFT_DSI (fields on the bounding surface from the device)"""
        out = np.zeros((self.NumSurfaces, 8))
        _check(host_lib().scuffb200_get_dsipft_matrix(self.ptr, KN.ptr, float(np.real(Omega)), float(np.imag(Omega)), IF.ptr if IF is not None else None,
                                                       float(DSIRadius), int(DSIPoints), (DSIMesh or "").encode(), out.ctypes.data))
        return out

    def GetFields(self, IF, KN, Omega, X):
        """total (scattered + incident) E, H at the rows of X (NX x 3): NX x 6 complex"""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1, 3))
        out = np.zeros((X.shape[0], 6), dtype=np.complex128)
        _check(host_lib().scuffb200_get_fields(self.ptr, IF.ptr if IF is not None else None, KN.ptr if KN is not None else None,
                                                float(np.real(Omega)), float(np.imag(Omega)), X.shape[0], X.ctypes.data, out.ctypes.data))
        return out

    def GetDyadicGFs(self, Omega, X, M, ScatteringOnly=True):
        """scattering parts of the electric / magnetic dyadic Green's functions at the rows of X (NX x 3, or NX x 6 = destination, source);
        M = the LU-factorised BEM matrix.  Returns (GE, GM), each (NX, 3, 3) complex"""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64)); X = X.reshape(-1, X.shape[-1])
        out = np.zeros((X.shape[0], 18), dtype=np.complex128)
        _check(host_lib().scuffb200_get_dyadic_gfs(self.ptr, M.ptr, float(np.real(Omega)), float(np.imag(Omega)), X.shape[0], X.shape[1], X.ctypes.data,
                                                    int(ScatteringOnly), out.ctypes.data))
        return out[:, :9].reshape(-1, 3, 3), out[:, 9:].reshape(-1, 3, 3)

    def GetRFMatrix(self, Omega, X):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1, 3))
        p = host_lib().scuffb200_get_rf_matrix(self.ptr, float(np.real(Omega)), float(np.imag(Omega)), X.shape[0], X.ctypes.data)
        if not p:
            raise ScuffError(host_lib().scuffb200_last_error().decode())
        return HMatrix(p)

    def surface_origin(self, ns):
        out = np.zeros(3)
        host_lib().scuffb200_surface_origin(self.ptr, ns, out.ctypes.data)
        return out

    def __init__(self, geofile):
        L = host_lib()
        self.ptr = L.scuffb200_geometry_open(os.fsencode(geofile))
        if not self.ptr:
            raise ScuffError(L.scuffb200_last_error().decode())
        info = (C.c_int * 8)()
        L.scuffb200_geometry_info(self.ptr, info)
        (self.NumSurfaces, self.TotalBFs, self.NumRegions, self.LDim,
         self.TotalPanels, self.TotalEdges, self.AllSurfacesClosed, _) = list(info)

    def surface(self, ns):
        L = host_lib()
        si = (C.c_int * 10)()
        if L.scuffb200_surface_info(self.ptr, ns, si):
            raise ScuffError("invalid surface index")
        NV, NP, NE = si[0], si[1], si[2]
        V = np.zeros((NV, 3)); PVI = np.zeros((NP, 3), dtype=np.int32); PC = np.zeros((NP, 3))
        PR = np.zeros(NP); PA = np.zeros(NP); ET = np.zeros((max(NE, 1), 8), dtype=np.int32); EL = np.zeros(max(NE, 1))
        L.scuffb200_surface_arrays(self.ptr, ns, V.ctypes.data, PVI.ctypes.data, PC.ctypes.data, PR.ctypes.data,
                                   PA.ctypes.data, ET.ctypes.data, EL.ctypes.data)
        return dict(Label=L.scuffb200_surface_label(self.ptr, ns).decode(), NV=NV, NP=NP, NE=NE, NBF=si[3],
                    IsPEC=bool(si[4]), Regions=(si[5], si[6]), BFOffset=si[7], Mate=si[8], NumHalfRWGEdges=si[9],
                    V=V, PVI=PVI, PC=PC, PR=PR, PA=PA, ET=ET[:NE], EL=EL[:NE])

    def regions(self):
        L = host_lib()
        return [(L.scuffb200_region_label(self.ptr, nr).decode(), L.scuffb200_region_material(self.ptr, nr).decode())
                for nr in range(self.NumRegions)]

    def eps_mu(self, omega):
        out = np.zeros((self.NumRegions, 4))
        _check(host_lib().scuffb200_eps_mu(self.ptr, float(np.real(omega)), float(np.imag(omega)), out.ctypes.data))
        return out[:, 0] + 1j * out[:, 1], out[:, 2] + 1j * out[:, 3]

    def transform(self, line):
        _check(host_lib().scuffb200_transform(self.ptr, line.encode()))

    def untransform(self):
        _check(host_lib().scuffb200_untransform(self.ptr))

    def mark_geometry_dirty(self):
        host_lib().scuffb200_mark_geometry_dirty(C.c_void_p(self.ptr))

    def zero_region(self, nr, zero=True):
        _check(host_lib().scuffb200_zero_region(self.ptr, nr, int(zero)))

    def AllocateBEMMatrix(self, PureImagFreq=False):
        p = host_lib().scuffb200_matrix_alloc(self.ptr, int(PureImagFreq))
        if not p:
            raise ScuffError(host_lib().scuffb200_last_error().decode())
        return HMatrix(p)

    def AssembleBEMMatrix(self, Omega, kBloch=None, M=None, to_host=True):
        if M is None:
            M = self.AllocateBEMMatrix()
        kb = None
        if kBloch is not None:
            kb = (C.c_double * 2)(float(kBloch[0]), float(kBloch[1]))
        _check(host_lib().scuffb200_assemble(self.ptr, M.ptr, float(np.real(Omega)), float(np.imag(Omega)),
                                              kb, int(to_host)))
        return M

    def AssembleBEMMatrixColumns(self, Omega, M, c0, c1):
        """column-block assembly for multi-GPU: upper-triangle entries of global columns [c0,c1), resident in HBM, not mirrored"""
        _check(host_lib().scuffb200_assemble_columns(self.ptr, M.ptr, float(np.real(Omega)), float(np.imag(Omega)), int(c0), int(c1)))
        return M

    def AssembleBEMMatrixBlock(self, nsa, nsb, Omega, M, kBloch=None, RowOffset=0, ColOffset=0, ABMBCache=None):
        kb = None
        if kBloch is not None:
            kb = (C.c_double * 2)(float(kBloch[0]), float(kBloch[1]))
        _check(host_lib().scuffb200_assemble_block_acc(self.ptr, M.ptr, nsa, nsb, float(np.real(Omega)),
                                                        float(np.imag(Omega)), kb, RowOffset, ColOffset, ABMBCache))
        return M

    def AssembleBEMMatrixBlockGrad(self, nsa, nsb, Omega, M, GradM=None, TorqueAxes=None, dMdT=None):
        """AssembleBEMMatrixBlock with its GradM[3] / dMdT[NumTorqueAxes] arguments (entries may be None); TorqueAxes: list of 3-vectors"""
        L = host_lib()
        gp = (C.c_void_p * 3)(*[(h.ptr if h is not None else None) for h in (GradM or [None] * 3)])
        nt = len(TorqueAxes) if TorqueAxes else 0
        ax = (C.c_double * max(3 * nt, 1))(*[float(x) for a in (TorqueAxes or []) for x in a])
        tp = (C.c_void_p * 3)(*([(h.ptr if h is not None else None) for h in (dMdT or [])] + [None] * 3)[:3])
        L.scuffb200_assemble_block_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                                    C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_void_p)]
        _check(L.scuffb200_assemble_block_grad(self.ptr, M.ptr, nsa, nsb, float(np.real(Omega)), float(np.imag(Omega)),
                                               gp if GradM else None, nt, ax, tp if dMdT else None))
        return M

    def CreateABMBAccelerator(self, nsa, nsb, PureImagFreq=False):
        """kBloch-independent matrix-block cache of one surface pair (AssembleBEMMatrix.cc:260-318); None for compact geometries"""
        p = host_lib().scuffb200_create_abmb(self.ptr, nsa, nsb, int(PureImagFreq))
        if not p and self.LDim:
            raise ScuffError(host_lib().scuffb200_last_error().decode())
        return p

    def DestroyABMBAccelerator(self, acc):
        host_lib().scuffb200_destroy_abmb(self.ptr, acc)

    def AssembleInnerCellBlocks(self, nsa, nsb, Omega, kBloch, M, RowOffset=0, ColOffset=0):
        """step 1 of the periodic branch of AssembleBEMMatrixBlock (AssembleBEMMatrix.cc:465-519): neighbour-cell blocks
        with Bloch phases; the outer-cell (Ewald GBar) part is not included"""
        kb = (C.c_double * 2)(float(kBloch[0]), float(kBloch[1]))
        _check(host_lib().scuffb200_assemble_inner_cells(self.ptr, M.ptr, nsa, nsb, float(np.real(Omega)),
                                                          float(np.imag(Omega)), kb, RowOffset, ColOffset))
        return M

    def GetOverlap(self, ns, nea, neb):
        """RWGSurface::GetOverlap: integral of f_a . f_b over the shared panels"""
        return host_lib().scuffb200_get_overlap(self.ptr, ns, nea, neb)

    def surface_zeta(self, ns, omega, X):
        """value of the surface's SURFACE_IMPEDANCE expression at angular frequency omega (internal units) and point X, or None"""
        out = (C.c_double * 2)()
        x = (C.c_double * 3)(*[float(v) for v in X])
        rc = host_lib().scuffb200_surface_zeta(self.ptr, ns, float(np.real(omega)), float(np.imag(omega)), x, out)
        if rc == 1:
            return None
        _check(rc)
        return complex(out[0], out[1])

    def PrepareNearPlans(self):
        """pair plans + FIPPI records of every block AssembleBEMMatrix visits, without assembling"""
        _check(host_lib().scuffb200_prepare_plans(self.ptr))

    def ppi_counts(self):
        out = (C.c_ulonglong * 5)()
        host_lib().scuffb200_ppi_counts(self.ptr, out)
        return list(out)

    def close(self):
        if self.ptr and _host is not None:
            _host.scuffb200_geometry_close(self.ptr)
            self.ptr = None

    def __del__(self):
        self.close()


def fippi_set_partition(rank, world):
    """this context computes only the FIPPI records with index = rank (mod world) until reset with (0, 1)"""
    _cuda_check(cuda_lib().scuffcuda_fippi_set_partition(context(), int(rank), int(world)))


def fippi_host_records():
    """zero-copy numpy view [count, 36] of the context's FIPPI table on the host (valid until the table grows)"""
    L = cuda_lib()
    L.scuffcuda_fippi_host_records.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_long)]
    p = C.POINTER(C.c_double)(); n = C.c_long()
    _cuda_check(L.scuffcuda_fippi_host_records(context(), C.byref(p), C.byref(n)))
    if n.value == 0:
        return np.zeros((0, 36))
    return np.ctypeslib.as_array(p, shape=(n.value, 36))


def fippi_commit():
    """send the (completed) host table to the device"""
    _cuda_check(cuda_lib().scuffcuda_fippi_commit(context()))


def PreloadCache(fn):
    _check(host_lib().scuffb200_preload_cache(os.fsencode(fn)))


def StoreCache(fn):
    _check(host_lib().scuffb200_store_cache(os.fsencode(fn)))


def ClearCache():
    _check(host_lib().scuffb200_clear_cache())


def cache_records():
    return host_lib().scuffb200_cache_records()


def last_timing():
    t = (C.c_double * 5)()
    cuda_lib().scuffcuda_get_last_timing(context(), t)
    return list(t)


def pair_counts():
    c = (C.c_ulonglong * 2)()
    cuda_lib().scuffcuda_get_pair_counts(context(), c)
    return list(c)


def launch_count():
    n = C.c_ulonglong()
    cuda_lib().scuffcuda_get_launch_count(context(), C.byref(n))
    return n.value


def transfer_bytes():
    b = (C.c_ulonglong * 2)()
    cuda_lib().scuffcuda_get_transfer_bytes(context(), b)
    return list(b)


def _cuda_check(rc):
    if rc:
        raise ScuffError(cuda_lib().scuffcuda_last_error(context()).decode())


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def gbar_ewald(k, kBloch, lbv, R, exclude_inner=False, ldim=2):
    """GBarVDEwald (GBarVDEwald.cc:931-996) on the device at points R[n,3]: complex array [n,8]"""
    R = _d(R).reshape(-1, 3)
    out = np.zeros((len(R), 16))
    kk = _d([np.real(k), np.imag(k)]); kb = _d(kBloch); lb = _d(lbv).ravel()
    _cuda_check(cuda_lib().scuffcuda_gbar_ewald(context(), kk.ctypes.data, kb.ctypes.data, lb.ctypes.data, ldim, int(exclude_inner),
                                                len(R), R.ctypes.data, out.ctypes.data))
    return out[:, 0::2] + 1j * out[:, 1::2]


class GBarAccelerator:
    """device table of the periodic Green's function (CreateGBarAccelerator, GBarAccelerator.cc:339-547)"""

    def __init__(self, k, kBloch, lbv, rho_min, rho_max, rel_tol=1e-3, exclude_inner=True, force_full=False, nmax=0, ldim=2):
        kk = _d([np.real(k), np.imag(k)]); kb = _d(kBloch); lb = _d(lbv).ravel()
        h = C.c_int()
        _cuda_check(cuda_lib().scuffcuda_gbar_create(context(), kk.ctypes.data, kb.ctypes.data, lb.ctypes.data, ldim, float(rho_min), float(rho_max),
                                                     float(rel_tol), int(exclude_inner), int(force_full), int(nmax), C.byref(h)))
        self.handle = h.value
        n = (C.c_int * 3)(); xmin = (C.c_double * 3)(); dx = (C.c_double * 3)(); lmax = C.c_double(); full = C.c_int()
        _cuda_check(cuda_lib().scuffcuda_gbar_info(context(), self.handle, n, xmin, dx, C.byref(lmax), C.byref(full)))
        self.N, self.XMin, self.DX, self.LMax, self.full = list(n), list(xmin), list(dx), lmax.value, bool(full.value)

    def table(self):
        ncell = (self.N[0] - 1) * (self.N[1] - 1) * (self.N[2] - 1)
        out = np.zeros(max(ncell, 0) * 128)
        if ncell > 0:
            _cuda_check(cuda_lib().scuffcuda_gbar_get_table(context(), self.handle, out.ctypes.data))
        return out

    def eval(self, R):
        """GetGBar: complex array [n,4] = G, dG/dx, dG/dy, dG/dz"""
        R = _d(R).reshape(-1, 3)
        out = np.zeros((len(R), 8))
        _cuda_check(cuda_lib().scuffcuda_gbar_eval(context(), self.handle, len(R), R.ctypes.data, out.ctypes.data))
        return out[:, 0::2] + 1j * out[:, 1::2]

    def close(self):
        if self.handle:
            cuda_lib().scuffcuda_gbar_destroy(context(), self.handle)
            self.handle = 0


def measure_fp64_peak():
    v = C.c_double()
    if cuda_lib().scuffcuda_measure_fp64_peak(context(), C.byref(v)):
        raise ScuffError(cuda_lib().scuffcuda_last_error(context()).decode())
    return v.value
