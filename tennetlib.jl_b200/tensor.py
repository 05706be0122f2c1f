"""Host-side index / tensor types and device handles.

`Index` and `HostTensor` mirror what the Julia shim extracts from an ITensor:
per index the (QN, dim) sectors and arrow; per tensor a dict {0-based sector coords -> dense block}.
`DeviceTensor` wraps a `tnl_tensor_t` living in HBM.
"""
from __future__ import annotations

import ctypes as C
import itertools
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, tnl_index_t

_ids = itertools.count(1 << 20)      # apart from the ids of foreign (test oracle / caller) index objects


class Index:
    __slots__ = ("id", "qns", "dims", "dir", "tags", "plev")

    def __init__(self, qns, dims, dir=+1, tags="", plev=0, id=None):
        self.qns = tuple(tuple(int(x) for x in q) for q in qns)
        self.dims = tuple(int(d) for d in dims)
        self.dir = int(dir)
        self.tags = tags
        self.plev = plev
        self.id = next(_ids) if id is None else id

    @property
    def dim(self):
        return sum(self.dims)

    @property
    def nsect(self):
        return len(self.dims)

    def copy(self, **kw):
        d = dict(qns=self.qns, dims=self.dims, dir=self.dir, tags=self.tags, plev=self.plev, id=self.id)
        d.update(kw)
        return Index(**d)

    def dag(self):
        return self.copy(dir=-self.dir)

    def prime(self, n=1):
        return self.copy(plev=self.plev + n)

    def __repr__(self):
        return f"Index({self.tags}|dir={self.dir:+d}|" + ",".join(f"{q}:{d}" for q, d in zip(self.qns, self.dims)) + ")"


class HostTensor:
    """Block-sparse host tensor (any object with `.inds` and `.blocks` is accepted wherever this is)."""

    def __init__(self, inds: Sequence[Index], blocks: Dict[Tuple[int, ...], np.ndarray] | None = None):
        self.inds = list(inds)
        self.blocks = {} if blocks is None else blocks

    def is_complex(self) -> bool:
        return any(np.iscomplexobj(b) for b in self.blocks.values())

    def to_dense(self):
        out = np.zeros([ix.dim for ix in self.inds], dtype=np.complex128 if self.is_complex() else np.float64)
        offs = [np.concatenate([[0], np.cumsum(ix.dims)]) for ix in self.inds]
        for c, b in self.blocks.items():
            out[tuple(slice(o[k], o[k + 1]) for o, k in zip(offs, c))] = b
        return out


def _index_array(inds):
    """ctypes array of tnl_index_t (+ keep-alive list)."""
    nq = len(inds[0].qns[0])
    arr = (tnl_index_t * len(inds))()
    keep = []
    for k, ix in enumerate(inds):
        dims = np.ascontiguousarray(ix.dims, dtype=np.int32)
        qns = np.ascontiguousarray(np.array(ix.qns, dtype=np.int32).reshape(-1))
        keep += [dims, qns]
        arr[k].nsect = len(ix.dims)
        arr[k].dir = ix.dir
        arr[k].dims = dims.ctypes.data_as(C.POINTER(C.c_int32))
        arr[k].qns = qns.ctypes.data_as(C.POINTER(C.c_int32))
    return arr, nq, keep


def flatten_blocks(t, dtype=np.float64):
    """NDTensors flat layout: (coords[nb,rank] i32, offsets[nb] i64, data), blocks column-major; `dtype`
    complex128 gives the interleaved (re, im) ComplexF64 storage, offsets counting complex elements."""
    keys = list(t.blocks.keys())
    rank = len(t.inds)
    coords = np.zeros((max(len(keys), 1), rank), dtype=np.int32)
    offsets = np.zeros(max(len(keys), 1), dtype=np.int64)
    chunks, off = [], 0
    for n, c in enumerate(keys):
        coords[n] = c
        offsets[n] = off
        b = np.asarray(t.blocks[c])
        if np.iscomplexobj(b) and np.dtype(dtype).kind != "c":
            if np.abs(b.imag).max(initial=0.0) > 0.0:
                raise NotImplementedError("complex site operators (complex MPO / CouplingModel tensors) are not "
                                          "supported on the device: only the state and the environments may be ComplexF64")
            b = b.real
        b = np.asarray(b, dtype=dtype)
        chunks.append(b.reshape(-1, order="F"))
        off += b.size
    data = np.ascontiguousarray(np.concatenate(chunks)) if chunks else np.zeros(1, dtype=dtype)
    return coords[:len(keys)] if keys else coords[:0], offsets[:len(keys)], data, len(keys)


class Context:
    """One CUDA context/stream of the library (`tnl_ctx_t`)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.tnl_ctx_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.tnl_ctx_destroy(self.h)
            self.h = None

    def counters(self):
        out = (C.c_double * 8)()
        check(self.lib.tnl_get_counters(self.h, out), self.h)
        k = ("gemm_flops", "transform_flops", "vec_bytes", "transform_bytes", "launches", "gemm_launches", "applies", "_")
        return dict(zip(k, list(out)))

    def reset_counters(self):
        check(self.lib.tnl_reset_counters(self.h), self.h)

    def sync(self):
        check(self.lib.tnl_ctx_sync(self.h), self.h)

    def reserve(self, nbytes: int):
        """Pre-grow the device memory pool (see tnl_ctx_reserve)."""
        check(self.lib.tnl_ctx_reserve(self.h, int(nbytes)), self.h)

    # ---- multi-GPU (one process per GPU)
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(self.lib.tnl_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world: int):
        check(self.lib.tnl_comm_init(self.h, uid, rank, world), self.h)

    def comm_set_sharding(self, enable: bool):
        """False: replicated, collective-free computation on every rank (parity reference of the sharded path)."""
        check(self.lib.tnl_comm_set_sharding(self.h, 1 if enable else 0), self.h)

    def comm_bench(self, n: int, reps: int, kind: int) -> float:
        ms = C.c_double()
        check(self.lib.tnl_comm_bench(self.h, n, reps, kind, C.byref(ms)), self.h)
        return ms.value

    def comm_destroy(self):
        check(self.lib.tnl_comm_destroy(self.h), self.h)

    def gemm_selftest(self, M, N, K, transA=False, transB=False, reps=1, verify=True):
        ms, err = C.c_double(), C.c_double()
        check(self.lib.tnl_gemm_selftest(self.h, M, N, K, int(transA), int(transB), reps, int(verify),
                                         C.byref(ms), C.byref(err)), self.h)
        return ms.value, err.value

    def profile_gemm(self, enable: bool):
        check(self.lib.tnl_profile_gemm(self.h, 1 if enable else 0), self.h)

    def profile_read(self):
        ms, n, fl, mx = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
        check(self.lib.tnl_profile_read(self.h, C.byref(ms), C.byref(n), C.byref(fl), C.byref(mx)), self.h)
        cat = (C.c_double * 8)()
        check(self.lib.tnl_profile_categories(self.h, cat), self.h)
        names = ("gemm", "transform", "vector", "collective")
        coll = (C.c_double * 12)()
        check(self.lib.tnl_profile_collectives(self.h, coll), self.h)
        kinds = ("allreduce_scalar", "reduce_scatter", "allgather", "allreduce_big", "fused_wait_consumed", "fused_slot_sum")
        return dict(total_ms=ms.value, launches=n.value, flops=fl.value, max_tflops=mx.value,
                    collective_ms={k: coll[i] for i, k in enumerate(kinds)},
                    collective_calls={k: int(coll[6 + i]) for i, k in enumerate(kinds)},
                    category_ms={k: cat[i] for i, k in enumerate(names)},
                    category_launches={k: int(cat[4 + i]) for i, k in enumerate(names[:3])},
                    host_plan_ms_total=cat[7])

    def timer_start(self, slot: int = 0):
        check(self.lib.tnl_timer_start(self.h, slot), self.h)

    def timer_stop(self, slot: int = 0) -> float:
        """CUDA-event milliseconds since timer_start(slot) on the library stream (synchronises)."""
        ms = C.c_double()
        check(self.lib.tnl_timer_stop(self.h, slot, C.byref(ms)), self.h)
        return ms.value


class DeviceTensor:
    def __init__(self, ctx: Context, handle, inds: Sequence[Index] | None = None):
        self.ctx = ctx
        self.h = handle
        self._inds = list(inds) if inds is not None else None

    # ---- construction
    @staticmethod
    def from_host(ctx: Context, t, nrow: int = 1) -> "DeviceTensor":
        arr, nq, keep = _index_array(t.inds)
        cplx = any(np.iscomplexobj(b) for b in t.blocks.values())
        coords, offsets, data, nb = flatten_blocks(t, np.complex128 if cplx else np.float64)
        h = C.c_void_p()
        fn = ctx.lib.tnl_tensor_import_c128 if cplx else ctx.lib.tnl_tensor_import
        check(fn(ctx.h, len(t.inds), nq, arr, nb, coords.ctypes.data, offsets.ctypes.data, data.ctypes.data, nrow,
                 C.byref(h)), ctx.h)
        return DeviceTensor(ctx, h, t.inds)

    @staticmethod
    def zeros(ctx: Context, inds, nrow: int = 1) -> "DeviceTensor":
        arr, nq, keep = _index_array(inds)
        h = C.c_void_p()
        check(ctx.lib.tnl_tensor_create(ctx.h, len(inds), nq, arr, nrow, C.byref(h)), ctx.h)
        return DeviceTensor(ctx, h, inds)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.tnl_tensor_free(self.h)
        except Exception:
            pass
        self.h = None

    # ---- queries
    @property
    def inds(self) -> List[Index]:
        """Index structure as the device sees it (ids are host-side only and regenerated when unknown)."""
        lib = self.ctx.lib
        rank, nq = C.c_int32(), C.c_int32()
        check(lib.tnl_tensor_rank(self.h, C.byref(rank), C.byref(nq)), self.ctx.h)
        out = []
        for k in range(rank.value):
            ns, dr = C.c_int32(), C.c_int32()
            check(lib.tnl_tensor_index(self.h, k, C.byref(ns), C.byref(dr), None, None, 0), self.ctx.h)
            dims = np.zeros(ns.value, dtype=np.int32)
            qns = np.zeros(ns.value * nq.value, dtype=np.int32)
            check(lib.tnl_tensor_index(self.h, k, C.byref(ns), C.byref(dr), dims.ctypes.data, qns.ctypes.data, ns.value),
                  self.ctx.h)
            qn = [tuple(int(x) for x in qns[s * nq.value:(s + 1) * nq.value]) for s in range(ns.value)]
            known = self._inds[k] if self._inds is not None and k < len(self._inds) else None
            if known is not None and known.qns == tuple(qn) and known.dims == tuple(int(d) for d in dims):
                out.append(known.copy(dir=dr.value))
            else:
                out.append(Index(qn, dims, dir=dr.value))
        return out

    def is_complex(self) -> bool:
        out = C.c_int32()
        check(self.ctx.lib.tnl_tensor_is_complex(self.h, C.byref(out)), self.ctx.h)
        return bool(out.value)

    def promote_(self):
        """real -> ComplexF64 (zero imaginary part), in place"""
        check(self.ctx.lib.tnl_tensor_promote(self.h), self.ctx.h)
        return self

    def to_host(self, drop_zero_blocks: bool = False) -> HostTensor:
        lib = self.ctx.lib
        nb, ne = C.c_int64(), C.c_int64()
        check(lib.tnl_tensor_export_size(self.h, C.byref(nb), C.byref(ne)), self.ctx.h)
        inds = self.inds
        rank = len(inds)
        coords = np.zeros((max(nb.value, 1), rank), dtype=np.int32)
        offsets = np.zeros(max(nb.value, 1), dtype=np.int64)
        data = np.zeros(max(ne.value, 1), dtype=np.complex128 if self.is_complex() else np.float64)
        check(lib.tnl_tensor_export(self.h, coords.ctypes.data, offsets.ctypes.data, data.ctypes.data), self.ctx.h)
        t = HostTensor(inds)
        for n in range(nb.value):
            c = tuple(int(x) for x in coords[n])
            shp = tuple(ix.dims[k] for ix, k in zip(inds, c))
            sz = int(np.prod(shp))
            blk = data[offsets[n]:offsets[n] + sz].reshape(shp, order="F").copy()
            if drop_zero_blocks and not blk.any():
                continue
            t.blocks[c] = blk
        return t

    def copy(self) -> "DeviceTensor":
        h = C.c_void_p()
        check(self.ctx.lib.tnl_tensor_copy(self.h, C.byref(h)), self.ctx.h)
        return DeviceTensor(self.ctx, h, self._inds)

    def fill_random(self, seed: int):
        check(self.ctx.lib.tnl_tensor_fill_random(self.h, C.c_uint64(seed)), self.ctx.h)
        return self

    def scale_index_(self, which: int, values):
        """T(..., i, ...) *= values[i] along index `which` (contraction with a diagonal matrix), in place."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        check(self.ctx.lib.tnl_tensor_scale_index(self.h, int(which), v.ctypes.data), self.ctx.h)
        return self

    # ---- VectorInterface
    def norm(self) -> float:
        out = C.c_double()
        check(self.ctx.lib.tnl_vec_norm(self.h, C.byref(out)), self.ctx.h)
        return out.value

    def dot(self, other: "DeviceTensor"):
        """<self, other> = sum conj(self) other; a complex number for complex tensors."""
        re, im = C.c_double(), C.c_double()
        check(self.ctx.lib.tnl_vec_dot_c(self.h, other.h, C.byref(re), C.byref(im)), self.ctx.h)
        return complex(re.value, im.value) if self.is_complex() else re.value

    def scale_(self, a: float):
        check(self.ctx.lib.tnl_vec_scale(self.h, float(a)), self.ctx.h)
        return self

    def axpy_(self, x: "DeviceTensor", a: float):
        check(self.ctx.lib.tnl_vec_axpy(self.h, x.h, float(a)), self.ctx.h)
        return self
