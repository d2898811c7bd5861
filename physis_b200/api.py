"""ctypes mirror of include/physis/physis_b200.h (host side, Python).

Names, argument order and error behaviour follow the Physis runtime API
(reference: include/physis/physis_common.h:78-103, physis_cuda.h:131-150,272-282):
errors are fatal in the native library (print + exit), grids are opaque
handles, Copyin/Copyout are synchronous.  Nothing here computes; every call
goes through the C ABI.
"""
import ctypes as C
import numpy as np

from ._lib import load_runtime

PS_MAX_DIM = 3
PS_INT, PS_LONG, PS_FLOAT, PS_DOUBLE, PS_USER = 0, 1, 2, 3, 4
PS_MAX, PS_MIN, PS_SUM, PS_PROD = 0, 1, 2, 3

KIND_GENERIC, KIND_DIFFUSION7_CLAMP, KIND_HIMENO19, KIND_HIMENO19_GOSA, KIND_PERIODIC7_STAGGERED = range(5)
MAX_GRIDS, MAX_SCALARS = 16, 8

_NP = {PS_INT: np.int32, PS_LONG: np.int64, PS_FLOAT: np.float32, PS_DOUBLE: np.float64}


class PSDomain(C.Structure):
    _fields_ = [("min", C.c_int32 * 3), ("max", C.c_int32 * 3),
                ("local_min", C.c_int32 * 3), ("local_max", C.c_int32 * 3)]


class MemberInfo(C.Structure):
    _fields_ = [("type", C.c_int), ("size", C.c_int), ("rank", C.c_int), ("dim", C.c_int * 5)]


class TypeInfo(C.Structure):
    _fields_ = [("type", C.c_int), ("size", C.c_int), ("num_members", C.c_int),
                ("members", C.POINTER(MemberInfo))]


class PSGridStruct(C.Structure):
    _fields_ = [("p", C.c_void_p), ("dim", C.c_int * 3), ("elm_size", C.c_int),
                ("num_dims", C.c_int), ("num_elms", C.c_int64), ("dev", C.c_void_p)]


class StencilDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("elm_type", C.c_int), ("dom", PSDomain),
                ("num_grids", C.c_int), ("grids", C.c_void_p * MAX_GRIDS),
                ("members", C.c_int * MAX_GRIDS), ("num_scalars", C.c_int),
                ("scalars", C.c_double * MAX_SCALARS), ("stencil", C.c_void_p),
                ("launch", C.c_void_p), ("name", C.c_char_p), ("written_mask", C.c_uint),
                ("z_reach", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("halo_bytes", C.c_uint64),
                ("last_kernel_ms", C.c_float), ("fused_pairs", C.c_uint64),
                ("fused_pairs_timed", C.c_uint64), ("fused_pair_ms", C.c_double),
                ("reduces_from_partials", C.c_uint64), ("plan_cache_hits", C.c_uint64),
                ("halo_wait_ns_sum", C.c_uint64), ("halo_wait_ns_max", C.c_uint64),
                ("halo_wait_ctas", C.c_uint64), ("halo_wait_launches", C.c_uint64),
                ("autotune_trials", C.c_uint64), ("autotuned_runs", C.c_uint64)]


# every extern "C" symbol include/physis/physis_b200.h declares
EXPORTED_SYMBOLS = [
    "PSInit", "PSFinalize", "PSDomain1DNew", "PSDomain2DNew", "PSDomain3DNew",
    "__PSGridNew", "__PSGridFree", "__PSGridCopyin", "__PSGridCopyout", "__PSGridSet",
    "__PSGridSwap", "__PSGridGetID", "__PSCheckCudaError", "PSGridCopyin", "PSGridCopyout",
    "PSGridFree", "__PSReduceGridFloat", "__PSReduceGridDouble", "__PSReduceGridInt",
    "__PSReduceGridLong", "__PSB200StencilRun", "__PSB200FusedPassCount", "__PSB200GetStream", "__PSB200Synchronize",
    "__PSB200TimerStart", "__PSB200TimerStopMs", "__PSB200GetStats", "__PSB200ResetStats",
    "__PSB200SetOption", "__PSB200LastTuning", "__PSB200Version", "__PSB200HostAlloc", "__PSB200HostFree", "__ps_trace",
    "__PSB200Rank", "__PSB200WorldSize", "__PSB200GridLocalSize", "__PSB200GridCopyinLocal",
    "__PSB200GridCopyoutLocal", "__PSB200Partition", "__PSB200GroupSelfTest",
]

_bound = False


def rt():
    """The runtime library with argtypes/restypes set."""
    global _bound
    lib = load_runtime()
    if _bound:
        return lib
    lib.PSInit.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_int]
    lib.PSInit.restype = None
    lib.PSDomain3DNew.argtypes = [C.c_int32] * 6
    lib.PSDomain3DNew.restype = PSDomain
    lib.PSDomain2DNew.argtypes = [C.c_int32] * 4
    lib.PSDomain2DNew.restype = PSDomain
    lib.PSDomain1DNew.argtypes = [C.c_int32] * 2
    lib.PSDomain1DNew.restype = PSDomain
    lib.__PSGridNew.argtypes = [C.POINTER(TypeInfo), C.c_int, C.POINTER(C.c_int), C.c_void_p]
    lib.__PSGridNew.restype = C.POINTER(PSGridStruct)
    lib.__PSGridFree.argtypes = [C.c_void_p, C.c_void_p]
    lib.__PSGridCopyin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.__PSGridCopyout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.PSGridCopyin.argtypes = [C.c_void_p, C.c_void_p]
    lib.PSGridCopyout.argtypes = [C.c_void_p, C.c_void_p]
    lib.PSGridFree.argtypes = [C.c_void_p]
    lib.__PSGridGetID.argtypes = [C.c_void_p]
    for n in ("Float", "Double", "Int", "Long"):
        f = getattr(lib, "__PSReduceGrid" + n)
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        f.restype = None
    lib.__PSB200StencilRun.argtypes = [C.c_int, C.c_int, C.POINTER(StencilDesc)]
    lib.__PSB200StencilRun.restype = C.c_float
    lib.__PSB200TimerStopMs.restype = C.c_float
    lib.__PSB200GetStats.argtypes = [C.POINTER(Stats)]
    lib.__PSB200SetOption.argtypes = [C.c_char_p]
    lib.__PSB200SetOption.restype = C.c_int
    lib.__PSB200Version.restype = C.c_char_p
    lib.__PSB200LastTuning.restype = C.c_char_p
    lib.__PSB200HostAlloc.argtypes = [C.c_size_t]
    lib.__PSB200HostAlloc.restype = C.c_void_p
    lib.__PSB200HostFree.argtypes = [C.c_void_p]
    lib.__PSB200GetStream.restype = C.c_void_p
    _bound = True
    return lib


# ---- thin API in the reference's vocabulary ---------------------------------

def PSInit(argv=None, num_dims=3, dims=(0, 0, 0)):
    argv = list(argv or ["physis"])
    argc = C.c_int(len(argv))
    arr = (C.c_char_p * (len(argv) + 1))(*[a.encode() for a in argv], None)
    parr = C.pointer(arr)
    # varargs: the maximum grid extents
    rt().PSInit(C.byref(argc), C.cast(C.pointer(parr), C.c_void_p), num_dims,
                *[C.c_int(d) for d in dims[:num_dims]])
    return [arr[i].decode() for i in range(argc.value)]


def PSFinalize():
    rt().PSFinalize()


def PSDomain3DNew(x0, x1, y0, y1, z0, z1):
    return rt().PSDomain3DNew(x0, x1, y0, y1, z0, z1)


def _type_info(ptype, members=None):
    """members: list of (PS type, array dims tuple) for a user-defined struct."""
    if members is None:
        ti = TypeInfo(ptype, np.dtype(_NP[ptype]).itemsize, 0, None)
        return ti, None
    arr = (MemberInfo * len(members))()
    off = 0
    maxal = 1
    for i, (t, adims) in enumerate(members):
        sz = np.dtype(_NP[t]).itemsize
        arr[i].type, arr[i].size, arr[i].rank = t, sz, len(adims)
        cnt = 1
        for k, d in enumerate(adims):
            arr[i].dim[k] = d
            cnt *= d
        off = (off + sz - 1) // sz * sz + sz * cnt
        maxal = max(maxal, sz)
    size = (off + maxal - 1) // maxal * maxal
    ti = TypeInfo(PS_USER, size, len(members), arr)
    return ti, arr


class Grid:
    """Handle wrapper: PSGrid{1,2,3}D<T>New / DeclareGrid user types."""

    def __init__(self, dims, ptype=PS_FLOAT, members=None):
        self.dims = tuple(int(d) for d in dims)
        self.ptype = PS_USER if members else ptype
        ti, self._keep = _type_info(ptype, members)
        self.elm_size = ti.size
        d = (C.c_int * 3)(*(list(self.dims) + [0] * (3 - len(self.dims))))
        self.h = getattr(rt(), "__PSGridNew")(C.byref(ti), len(self.dims), d, None)  # no name mangling
        if not self.h:
            raise MemoryError("__PSGridNew returned INVALID_GRID")
        self.num_elms = int(np.prod(self.dims))

    @property
    def ptr(self):
        return C.cast(self.h, C.c_void_p)

    def copyin(self, host):
        host = np.ascontiguousarray(host)
        assert host.nbytes == self.elm_size * self.num_elms, (host.nbytes, self.elm_size, self.num_elms)
        rt().PSGridCopyin(self.ptr, host.ctypes.data)

    def copyout(self, dtype=None):
        if dtype is None:
            dtype = _NP[self.ptype] if self.ptype != PS_USER else np.uint8
        out = np.empty(self.elm_size * self.num_elms // np.dtype(dtype).itemsize, dtype=dtype)
        rt().PSGridCopyout(self.ptr, out.ctypes.data)
        return out

    def reduce(self, op):
        name = {PS_FLOAT: "Float", PS_DOUBLE: "Double", PS_INT: "Int", PS_LONG: "Long"}[self.ptype]
        out = np.zeros(1, dtype=_NP[self.ptype])
        getattr(rt(), "__PSReduceGrid" + name)(out.ctypes.data, op, self.ptr)
        return out[0]

    def set(self, index, value_bytes):
        buf = np.frombuffer(bytes(value_bytes), dtype=np.uint8).copy()
        getattr(rt(), "__PSGridSet")(self.ptr, C.c_void_p(buf.ctypes.data), *[C.c_int(i) for i in index])

    def free(self):
        if self.h:
            rt().PSGridFree(self.ptr)
            self.h = None


def stencil_desc(kind, dom, grids, scalars=(), members=None, elm_type=PS_FLOAT, name=b"sweep"):
    d = StencilDesc()
    d.kind = kind
    d.elm_type = elm_type
    d.dom = dom
    d.num_grids = len(grids)
    for i, g in enumerate(grids):
        d.grids[i] = C.cast(g.h, C.c_void_p).value
        d.members[i] = -1 if members is None else members[i]
    d.num_scalars = len(scalars)
    for i, s in enumerate(scalars):
        d.scalars[i] = float(s)
    d.name = name
    return d


def stencil_run(iters, descs):
    arr = (StencilDesc * len(descs))(*descs)
    return rt().__PSB200StencilRun(iters, len(descs), arr)


def set_option(kv):
    if rt().__PSB200SetOption(kv.encode()) != 0:
        raise ValueError(f"unknown physis_b200 option {kv!r}")


def last_tuning():
    """What option autotune=1 last settled on (``__PSB200LastTuning``)."""
    return rt().__PSB200LastTuning().decode()


def stats():
    s = Stats()
    rt().__PSB200GetStats(C.byref(s))
    return s


def pinned_empty(nbytes, dtype=np.uint8):
    """Page-locked host array (for end-to-end timing with DMA-able buffers)."""
    p = rt().__PSB200HostAlloc(nbytes)
    buf = (C.c_uint8 * nbytes).from_address(p)
    a = np.frombuffer(buf, dtype=dtype)
    return a, p
