"""ctypes binding of lib/libhifiles_b200.so: the C API of the host mirror (host/capi.cpp) and the device C ABI
(include/hifiles_b200.h).  This module holds no numerics: every call goes into the shared library, which has no CPU
fallback (creating a run without a CUDA device raises).  The directory name carries a hyphen, so import it with
`hifiles_b200 = load_package()` from __graft_entry__ / tests.conftest (importlib by path)."""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libhifiles_b200.so")
DATA_DIR = os.path.join(PKG_DIR, "data")

HF_ARRAY_IDS = dict(
    disu_upts=0, disu_upts1=1, div_tconf_upts=2, disu_fpts=3, tdisf_upts=4, norm_tdisf_fpts=5, norm_tconf_fpts=6,
    delta_disu_fpts=7, grad_disu_upts=8, grad_disu_fpts=9, src_upts=10, dt_local=11, sensor=12, sgsf_upts=13, sgsf_fpts=14, disuf_upts=15, Lu=16, Le=17)
ELE_TYPES = dict(tri=0, quad=1, tet=2, pri=3, hex=4)
ELES_OPS = dict(extrapolate_solution=0, calculate_gradient=1, evaluate_invFlux=2, correct_gradient=3, evaluate_viscFlux=4,
                extrapolate_totalFlux=5, calculate_divergence=6, calculate_corrected_divergence=7, evaluate_invFlux_over_int=8,
                shock_capture=9, extrapolate_sgsFlux=10, calc_sgs_terms=11)


class HiFiLESError(RuntimeError):
    pass


def build(verbose=False):
    """Compile the library in-tree with the package Makefile (nvcc, sm_100a)."""
    r = subprocess.run(["make", "-C", PKG_DIR, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise HiFiLESError("building libhifiles_b200.so failed")


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HiFiLESError("%s is missing: run __graft_entry__.build() (no CPU fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.hifiles_last_error.restype = C.c_char_p
    L.hf_dev_last_error.restype = C.c_char_p
    L.hifiles_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_longlong, C.c_int, C.POINTER(C.c_void_p)]
    L.hifiles_destroy.argtypes = [C.c_void_p]
    L.hifiles_device_ctx.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.hifiles_calc_residual.argtypes = [C.c_void_p, C.c_int]
    L.hifiles_advance_solution.argtypes = [C.c_void_p, C.c_int]
    L.hifiles_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hifiles_norm_residual.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    L.hifiles_copy_solution_to_host.argtypes = [C.c_void_p]
    L.hifiles_get_scalar.argtypes = [C.c_void_p, C.c_char_p]
    L.hifiles_get_scalar.restype = C.c_double
    L.hifiles_get_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
    L.hifiles_n_inters.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hifiles_n_eles.argtypes = [C.c_void_p, C.c_int]
    L.hifiles_calc_time_step.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.hifiles_nccl_init.argtypes = [C.c_void_p, C.c_char_p]
    L.hifiles_write_vtu.argtypes = [C.c_void_p, C.c_int]
    L.hifiles_write_restart.argtypes = [C.c_void_p, C.c_int]
    L.hf_dev_nccl_unique_id.argtypes = [C.c_char_p]
    L.hf_dev_eles_op.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hf_dev_int_inters_op.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hf_dev_bdy_inters_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    L.hf_dev_mpi_inters_op.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hf_dev_calc_residual.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.hf_dev_advance_solution.argtypes = [C.c_void_p, C.c_int]
    L.hf_dev_rk_stage.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
    L.hf_dev_run_steps.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.hf_dev_check_residual.argtypes = [C.c_void_p]
    L.hf_dev_download.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.hf_dev_upload.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.hf_dev_upload_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.hf_dev_upload_commit.argtypes = [C.c_void_p, C.c_int]
    L.hf_dev_residual_norm.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.hf_dev_sync.argtypes = [C.c_void_p]
    L.hf_dev_launch_count.argtypes = [C.c_void_p]
    L.hf_dev_launch_count.restype = C.c_longlong
    L.hf_dev_set_mode.argtypes = [C.c_void_p, C.c_int]
    L.hf_dev_set_dt.argtypes = [C.c_void_p, C.c_double]
    L.hf_dev_timer_start.argtypes = [C.c_void_p]
    L.hf_dev_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.hf_dev_fused_status.argtypes = [C.c_void_p]
    L.hf_dev_fused_status.restype = C.c_char_p
    L.hf_dev_elem_status.argtypes = [C.c_void_p]
    L.hf_dev_elem_status.restype = C.c_char_p
    L.hf_dev_fused_variant.argtypes = [C.c_void_p]
    L.hf_dev_fused_variant.restype = C.c_char_p
    L.hf_dev_kernel_timer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    _lib = L
    return L


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    if lib().hf_dev_nccl_unique_id(buf) != 0:
        raise HiFiLESError(lib().hf_dev_last_error().decode())
    return buf.raw


class Run:
    """One HiFiLES run = what the reference's main() owns (struct solution + run_input): mesh, elements, interfaces on
    the host; solution and all per-stage work on the device."""

    def __init__(self, input_file, rank=0, nproc=1, part=None, host_only=False, nccl_id=None):
        L = lib()
        os.environ.setdefault("HIFILES_HOME", PKG_DIR)
        self._h = C.c_void_p()
        part_p, n_part = None, 0
        if part is not None:
            self._part = np.ascontiguousarray(part, dtype=np.int32)
            part_p = self._part.ctypes.data_as(C.POINTER(C.c_int))
            n_part = self._part.size
        flags = 1 if host_only else 0
        if L.hifiles_create(str(input_file).encode(), rank, nproc, part_p, n_part, flags, C.byref(self._h)) != 0:
            raise HiFiLESError(L.hifiles_last_error().decode())
        self.host_only = host_only
        self.rank, self.nproc = rank, nproc
        if nproc > 1 and not host_only:
            if nccl_id is None:
                raise HiFiLESError("nproc > 1 needs the NCCL unique id of rank 0")
            self._ck(L.hifiles_nccl_init(self._h, nccl_id))

    def _ck(self, status):
        if status != 0:
            raise HiFiLESError(lib().hifiles_last_error().decode() or lib().hf_dev_last_error().decode())

    def _ckd(self, status):
        if status != 0:
            raise HiFiLESError(lib().hf_dev_last_error().decode())

    def close(self):
        if self._h:
            lib().hifiles_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- queries -----------------------------------------------------------------------------------------------
    def scalar(self, name):
        return lib().hifiles_get_scalar(self._h, name.encode())

    def n_eles(self, ele_type):
        return lib().hifiles_n_eles(self._h, ELE_TYPES[ele_type])

    def n_inters(self, kind, inter_type):
        return lib().hifiles_n_inters(self._h, dict(int=0, bdy=1, mpi=2)[kind], inter_type)

    def ele_types(self):
        return [t for t in ELE_TYPES if self.n_eles(t) > 0]

    def host_array(self, name):
        """Copy of a host-side array by the name oracle/ref_dump.cpp uses ('hex.opp_0', 'int_quad.idx_l', ...)."""
        p = C.c_void_p(); dt = C.c_int(); nd = C.c_int(); dims = (C.c_longlong * 4)()
        self._ck(lib().hifiles_get_array(self._h, name.encode(), C.byref(p), C.byref(dt), C.byref(nd), dims))
        shape = tuple(dims[i] for i in range(nd.value))
        n = int(np.prod(shape))
        ctype = C.c_double if dt.value == 0 else C.c_int
        if n == 0 or not p.value:
            return np.zeros(shape, dtype=np.float64 if dt.value == 0 else np.int32, order="F")
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).copy()
        return a.reshape(shape, order="F")

    # ---- the reference's call sequence --------------------------------------------------------------------------
    def calc_residual(self, rk_stage=0):
        self._ck(lib().hifiles_calc_residual(self._h, rk_stage))

    def advance_solution(self, rk_stage):
        self._ck(lib().hifiles_advance_solution(self._h, rk_stage))

    def calc_time_step(self):
        dt = C.c_double()
        self._ck(lib().hifiles_calc_time_step(self._h, C.byref(dt)))
        return dt.value

    def run(self, n_steps, fused=True):
        self._ck(lib().hifiles_run(self._h, n_steps, 1 if fused else 0))

    def norm_residual(self):
        n = int(self.scalar("n_dims")) + 2 if int(self.scalar("equation")) == 0 else 1
        out = (C.c_double * 8)()
        self._ck(lib().hifiles_norm_residual(self._h, out, n))
        return np.array(out[:n])

    # ---- device access --------------------------------------------------------------------------------------------
    @property
    def ctx(self):
        p = C.c_void_p()
        self._ck(lib().hifiles_device_ctx(self._h, C.byref(p)))
        return p

    def _dev_shape(self, ele_type, which):
        ne, nu, nf, nfl, nd, _ = [int(v) for v in self.host_array(ele_type + ".sizes")]
        return dict(
            disu_upts=(nu, ne, nfl), disu_upts1=(nu, ne, nfl), div_tconf_upts=(nu, ne, nfl), disu_fpts=(nf, ne, nfl),
            tdisf_upts=(nu, ne, nfl, nd), norm_tdisf_fpts=(nf, ne, nfl), norm_tconf_fpts=(nf, ne, nfl),
            delta_disu_fpts=(nf, ne, nfl), grad_disu_upts=(nu, ne, nfl, nd), grad_disu_fpts=(nf, ne, nfl, nd),
            src_upts=(nu, ne, nfl), dt_local=(ne,), sensor=(ne,), sgsf_upts=(nu, ne, nfl, nd), sgsf_fpts=(nf, ne, nfl, nd), disuf_upts=(nu, ne, nfl), Lu=(nu, ne, 3 if nd == 2 else 6), Le=(nu, ne, nd))[which]

    def download(self, ele_type, which):
        shape = self._dev_shape(ele_type, which)
        out = np.empty(int(np.prod(shape)), dtype=np.float64)
        self._ckd(lib().hf_dev_download(self.ctx, ELE_TYPES[ele_type], HF_ARRAY_IDS[which], out.ctypes.data, out.size))
        return out.reshape(shape, order="F")

    def upload(self, ele_type, which, arr):
        a = np.asfortranarray(arr, dtype=np.float64).ravel(order="F")
        self._ckd(lib().hf_dev_upload(self.ctx, ELE_TYPES[ele_type], HF_ARRAY_IDS[which], a.ctypes.data, a.size))

    def upload_begin(self, ele_type, which, arr):
        """hf_dev_upload_begin: returns the (page-locked) staging array that must stay alive until the commit has been consumed."""
        import torch
        a = np.asfortranarray(arr, dtype=np.float64).ravel(order="F")
        pinned = torch.empty(a.size, dtype=torch.float64, pin_memory=True)
        pinned.numpy()[:] = a
        self._ckd(lib().hf_dev_upload_begin(self.ctx, ELE_TYPES[ele_type], HF_ARRAY_IDS[which], C.c_void_p(pinned.data_ptr()), a.size))
        return pinned

    def upload_commit(self, ele_type):
        self._ckd(lib().hf_dev_upload_commit(self.ctx, ELE_TYPES[ele_type]))

    def eles_op(self, ele_type, op):
        self._ckd(lib().hf_dev_eles_op(self.ctx, ELE_TYPES[ele_type], ELES_OPS[op]))

    def int_inters_op(self, inter_type, op):
        self._ckd(lib().hf_dev_int_inters_op(self.ctx, inter_type, op))

    def bdy_inters_op(self, inter_type, op, time=0.0):
        self._ckd(lib().hf_dev_bdy_inters_op(self.ctx, inter_type, op, time))

    def set_mode(self, fused):
        self._ckd(lib().hf_dev_set_mode(self.ctx, 1 if fused else 0))

    def fused_status(self):
        return lib().hf_dev_fused_status(self.ctx).decode()

    def write_vtu(self, it):
        """output::write_vtu: Paraview file(s) of the current solution in the working directory"""
        self._ck(lib().hifiles_write_vtu(self._h, int(it)))

    def write_restart(self, it):
        """output::write_restart_ascii into the working directory (one file per rank in Rest_<iter>/ when partitioned)"""
        self._ck(lib().hifiles_write_restart(self._h, int(it)))

    def elem_status(self):
        return lib().hf_dev_elem_status(self.ctx).decode()

    def fused_variant(self):
        return lib().hf_dev_fused_variant(self.ctx).decode()

    def rk_stage(self, stage, time=0.0, keep_residual=False):
        self._ckd(lib().hf_dev_rk_stage(self.ctx, stage, time, 1 if keep_residual else 0))

    def run_steps(self, n_steps, time0=0.0):
        self._ckd(lib().hf_dev_run_steps(self.ctx, n_steps, time0))

    def sync(self):
        self._ckd(lib().hf_dev_sync(self.ctx))

    def launch_count(self):
        return lib().hf_dev_launch_count(self.ctx)

    def timer_start(self):
        self._ckd(lib().hf_dev_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_float()
        self._ckd(lib().hf_dev_timer_stop(self.ctx, C.byref(ms)))
        return ms.value

    def kernel_timer(self, enable=None):
        """Enable (True/False) per-kernel CUDA-event timing of the dominant fused kernel, or read (ms_total, launches)."""
        ms = C.c_double(); n = C.c_longlong()
        mode = -1 if enable is None else (1 if enable else 0)
        self._ckd(lib().hf_dev_kernel_timer(self.ctx, mode, C.byref(ms), C.byref(n)))
        return ms.value, n.value
