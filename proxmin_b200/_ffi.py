"""ctypes binding of libproxmin_b200.so (include/proxmin_b200.h).

There is no CPU fallback: if the shared library is missing or no B200 is visible,
the first device call raises.  Importing this module never touches the GPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libproxmin_b200.so")

PMX_MAX_OPS = 8
OP_ID, OP_ZERO, OP_PLUS, OP_UNITY, OP_MIN, OP_MAX, OP_HARD, OP_SOFT, OP_MAXENT, OP_MAXENT64 = range(10)
A, S, GA, GS, MA, MS, VA, VS, VHA, VHS = range(10)
SCHEMES = {"adam": 0, "nadam": 1, "amsgrad": 2, "padam": 3, "adamx": 4, "radam": 5}

EW_EXTRAP, EW_ADD, EW_SUB, EW_DX_ACC, EW_ZU, EW_DOT_DIFF, EW_MAXABS, EW_SUMSQ, EW_BB, EW_AXPY = range(10)
ERR_CUDA, ERR_ARG, ERR_NCCL, ERR_UNSUPPORTED, ERR_NONFINITE = -1, -2, -3, -4, -5


class ProxOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("relative", C.c_int32), ("axis", C.c_int32), ("thresh", C.c_float)]


class Prox(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("ops", ProxOp * PMX_MAX_OPS)]


class PgmOpts(C.Structure):
    _fields_ = [("prox_A", Prox), ("prox_S", Prox), ("accelerated", C.c_int32), ("e_rel_A", C.c_float),
                ("e_rel_S", C.c_float), ("kernel", C.c_int32), ("check_every", C.c_int32)]


class AdaproxOpts(C.Structure):
    _fields_ = [("prox_A", Prox), ("prox_S", Prox), ("has_prox_A", C.c_int32), ("has_prox_S", C.c_int32),
                ("scheme", C.c_int32), ("b2", C.c_double), ("eps", C.c_double), ("p", C.c_double),
                ("e_rel_A", C.c_float), ("e_rel_S", C.c_float), ("check_convergence", C.c_int32),
                ("prox_max_iter", C.c_int32), ("has_vhat", C.c_int32), ("kernel", C.c_int32),
                ("step_mode", C.c_int32), ("alpha_A", C.c_float), ("alpha_S", C.c_float)]


class BsdmmOpts(C.Structure):
    _fields_ = [("prox_A", Prox), ("prox_S", Prox), ("n_g_A", C.c_int32), ("n_g_S", C.c_int32),
                ("proxs_g_A", Prox * 4), ("proxs_g_S", Prox * 4), ("e_rel_A", C.c_float), ("e_rel_S", C.c_float),
                ("e_abs_A", C.c_float), ("e_abs_S", C.c_float), ("kernel", C.c_int32)]


class AdmmOpts(C.Structure):
    _fields_ = [("n_g", C.c_int32), ("proxs_g", Prox * 4), ("e_rel", C.c_float), ("e_abs", C.c_float),
                ("dual_uses_step_g", C.c_int32)]


class DeviceError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded shared library (raises if it has not been built: no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DeviceError(
                "libproxmin_b200.so is missing (%s). Build it with `python -m proxmin_b200.build`; "
                "there is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    pi, pf, pd = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.pmx_last_error.restype = C.c_char_p
    sig = {
        "pmx_version": [],
        "pmx_device_count": [pi],
        "pmx_ctx_create": [i32, C.POINTER(vp)],
        "pmx_ctx_destroy": [vp],
        "pmx_ctx_sync": [vp],
        "pmx_ctx_trim": [vp],
        "pmx_ctx_launch_count": [vp, C.POINTER(C.c_longlong)],
        "pmx_ctx_device_info": [vp, C.c_char_p, i32, pi, C.POINTER(sz)],
        "pmx_ctx_profile": [vp, i32],
        "pmx_ctx_profile_read": [vp, pf, pi],
        "pmx_comm_unique_id": [vp],
        "pmx_comm_init": [vp, vp, i32, i32],
        "pmx_comm_allreduce_sum": [vp, vp, sz],
        "pmx_comm_peer_enabled": [vp, pi],
        "pmx_malloc": [vp, sz, C.POINTER(vp)],
        "pmx_free": [vp, vp],
        "pmx_memset": [vp, vp, i32, sz],
        "pmx_h2d": [vp, vp, vp, sz],
        "pmx_d2h": [vp, vp, vp, sz],
        "pmx_d2d": [vp, vp, vp, sz],
        "pmx_host_alloc": [sz, C.POINTER(vp)],
        "pmx_host_free": [vp],
        "pmx_timer_start": [vp],
        "pmx_timer_stop": [vp, pf],
        "pmx_prox_apply": [vp, C.POINTER(Prox), vp, i32, i32, f32],
        "pmx_nmf_grad": [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32],
        "pmx_nmf_lipschitz": [vp, vp, vp, i32, i32, i32, pf, pf],
        "pmx_nmf_create": [vp, i32, i32, i32, C.POINTER(vp)],
        "pmx_nmf_destroy": [vp],
        "pmx_nmf_set_Y": [vp, vp, sz, i32, i32],
        "pmx_nmf_set_W": [vp, vp, sz, i32, i32],
        "pmx_nmf_gradient": [vp, pd],
        "pmx_nmf_set": [vp, i32, vp],
        "pmx_nmf_get": [vp, i32, vp],
        "pmx_nmf_device_ptr": [vp, i32, C.POINTER(vp)],
        "pmx_nmf_loss": [vp, pd],
        "pmx_nmf_pgm_begin": [vp, C.POINTER(PgmOpts)],
        "pmx_nmf_pgm_run": [vp, i32, pi, pi, pi, pf, pf],
        "pmx_nmf_adaprox_begin": [vp, C.POINTER(AdaproxOpts)],
        "pmx_nmf_adaprox_run": [vp, i32, pd, pd, pi, pi, pi, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
        "pmx_nmf_bsdmm_begin": [vp, C.POINTER(BsdmmOpts)],
        "pmx_nmf_bsdmm_run": [vp, i32, pi, pi, pi],
        "pmx_admm_create": [vp, sz, C.POINTER(AdmmOpts), C.POINTER(vp)],
        "pmx_admm_destroy": [vp],
        "pmx_admm_set": [vp, vp, vp],
        "pmx_admm_get": [vp, vp],
        "pmx_admm_init_zu": [vp],
        "pmx_admm_step": [vp, C.c_double, pi, pi, pd],
        "pmx_admm_run": [vp, C.c_double, i32, pi, pi, pd],
        "pmx_admm_stats": [vp, C.POINTER(C.c_longlong), pi],
        "pmx_pgm_update": [vp, C.POINTER(Prox), vp, vp, vp, i32, i32, f32, pd, pd],
        "pmx_axis_sum": [vp, vp, i32, i32, i32, pd],
        "pmx_matmul": [vp, vp, vp, vp, i32, i32, i32, i32],
        "pmx_ew": [vp, i32, sz, vp, vp, vp, vp, f32, f32, vp, vp, vp, pd],
        "pmx_adaprox_moments": [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, i32, f32, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_double, i32, pf],
        "pmx_adaprox_sub": [vp, C.POINTER(Prox), vp, vp, vp, vp, i32, i32, vp, i32, f32, f32, pd],
    }
    for name, args in sig.items():
        fn = getattr(L, name)  # AttributeError here = header and library disagree: fail loudly
        fn.argtypes = args
        fn.restype = C.c_int
    L._pmx_exports = sorted(sig)


EXPORTS = ["pmx_last_error"]  # filled with every declared symbol on first lib() call


def check(status):
    if status == 0:
        return
    msg = lib().pmx_last_error().decode("utf-8", "replace")
    if status == ERR_NONFINITE:
        raise np.linalg.LinAlgError(msg or "Array must not contain infs or NaNs")
    if status == ERR_ARG:
        raise ValueError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise DeviceError("proxmin_b200 error %d: %s" % (status, msg))


# ----------------------------------------------------------------------------- context

class Context:
    """One CUDA context + stream pair per process (one process per GPU)."""

    def __init__(self, device=None):
        L = lib()
        if device is None:
            device = int(os.environ.get("PROXMIN_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        n = C.c_int(0)
        check(L.pmx_device_count(C.byref(n)))
        if n.value < 1:
            raise DeviceError("no CUDA device visible; proxmin_b200 has no CPU fallback")
        self.handle = C.c_void_p()
        check(L.pmx_ctx_create(device % n.value, C.byref(self.handle)))
        self.device = device % n.value
        self.world, self.rank = 1, 0

    def sync(self):
        check(lib().pmx_ctx_sync(self.handle))

    def trim(self):
        """Give the device blocks cached from destroyed solver handles back to the driver."""
        check(lib().pmx_ctx_trim(self.handle))

    def launches(self):
        n = C.c_longlong(0)
        check(lib().pmx_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    def device_info(self):
        name = C.create_string_buffer(128)
        sm, mem = C.c_int(0), C.c_size_t(0)
        check(lib().pmx_ctx_device_info(self.handle, name, 128, C.byref(sm), C.byref(mem)))
        return name.value.decode(), sm.value, mem.value

    def profile(self, enable):
        check(lib().pmx_ctx_profile(self.handle, int(enable)))

    def profile_read(self):
        ms, n = C.c_float(0), C.c_int(0)
        check(lib().pmx_ctx_profile_read(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- memory
    def malloc(self, nbytes):
        p = C.c_void_p()
        check(lib().pmx_malloc(self.handle, nbytes, C.byref(p)))
        return p

    def free(self, p):
        if p:
            check(lib().pmx_free(self.handle, p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        check(lib().pmx_h2d(self.handle, dptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def d2h(self, arr, dptr):
        assert arr.flags.c_contiguous
        check(lib().pmx_d2h(self.handle, arr.ctypes.data_as(C.c_void_p), dptr, arr.nbytes))

    def upload(self, arr):
        a32 = np.ascontiguousarray(arr, dtype=np.float32)
        p = self.malloc(a32.nbytes)
        self.h2d(p, a32)
        return p

    def timer_start(self):
        check(lib().pmx_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float(0)
        check(lib().pmx_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    # -- multi-GPU
    def peer_enabled(self):
        """True when the sharded solvers exchange over CUDA-IPC peer memory (comm.cu), False = NCCL all-reduces"""
        v = C.c_int(0)
        check(lib().pmx_comm_peer_enabled(self.handle, C.byref(v)))
        return bool(v.value)

    def unique_id(self):
        buf = C.create_string_buffer(128)
        check(lib().pmx_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid, world, rank):
        buf = C.create_string_buffer(bytes(uid), 128)
        check(lib().pmx_comm_init(self.handle, buf, world, rank))
        self.world, self.rank = world, rank


_ctx = None


def context():
    global _ctx
    if _ctx is None:
        _ctx = Context()
    return _ctx


def make_prox(ops):
    """ops: list of (opcode, relative, axis, thresh) in application order -> Prox struct."""
    if len(ops) > PMX_MAX_OPS:
        raise NotImplementedError("proximal chain longer than %d primitive operators" % PMX_MAX_OPS)
    p = Prox()
    p.n_ops = len(ops)
    for i, (op, rel, axis, thr) in enumerate(ops):
        p.ops[i].op, p.ops[i].relative, p.ops[i].axis, p.ops[i].thresh = op, int(rel), int(axis), float(thr)
    return p
