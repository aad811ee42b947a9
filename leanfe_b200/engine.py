"""Thin host wrapper: PyTorch tensors as device buffers, every compute call goes to liblfe.so.

PyTorch is plumbing here (allocation, streams, H2D/D2H copies); no arithmetic of the hot path is
done with torch ops.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi as capi

_TORCH_TO_LFE = {
    torch.int32: capi.LFE_I32, torch.int64: capi.LFE_I64, torch.float64: capi.LFE_F64, torch.uint8: capi.LFE_U8,
}
if hasattr(torch, "uint32"):
    _TORCH_TO_LFE[torch.uint32] = capi.LFE_U32
if hasattr(torch, "uint64"):
    _TORCH_TO_LFE[torch.uint64] = capi.LFE_U64


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("leanfe_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


class Engine:
    """One liblfe handle on one CUDA device."""

    def __init__(self, device: int | None = None, profiling: bool = False):
        _require_cuda()
        self.lib = capi.load_library()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        capi.check(self.lib.lfe_create(self.device_index, C.byref(h)))
        self.h = h
        self.world, self.rank = 1, 0
        if profiling:
            self.set_profiling(True)

    def close(self):
        if getattr(self, "_uploader", None) is not None:
            self.lib.lfe_uploader_destroy(self._uploader)
            self._uploader = None
        if getattr(self, "h", None):
            self.lib.lfe_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def copy_stream(self):
        """One persistent side stream for host->device copies (overlapped with the id-only compute)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._copy_stream

    def host_uploader(self):
        """The native pageable-memory uploader of this engine (lfe_uploader_*: pinned ring + staging threads),
        created on first use: 8 x 32 MB pinned slots; LFE_UPLOAD_THREADS staging threads (default: a share of the
        host's cores, so that one rank per GPU does not oversubscribe the box)."""
        if getattr(self, "_uploader", None) is None:
            import os
            try:
                cores = len(os.sched_getaffinity(0))
            except Exception:
                cores = os.cpu_count() or 4
            ranks = max(torch.cuda.device_count() if self.world > 1 else 1, 1)
            # staging has to keep up with PCIe (55 GB/s): ~7 GB/s per thread -> up to 12 threads, leaving cores for
            # the caller and the dispatcher; one rank per GPU shares the host
            threads = int(os.environ.get("LFE_UPLOAD_THREADS", max(2, min(12, (cores * 3) // (4 * ranks) or 1))))
            u = C.c_void_p()
            capi.check(self.lib.lfe_uploader_create(self.h, 32 << 20, 8, threads, C.byref(u)))
            self._uploader = u
        return self._uploader

    def upload_enqueue(self, host_ptr: int, nbytes: int, dst: torch.Tensor, copy_stream) -> int:
        t = C.c_int64(0)
        capi.check(self.lib.lfe_uploader_enqueue(self.host_uploader(), C.c_void_p(host_ptr), int(nbytes), self._p(dst),
                                                 C.c_void_p(copy_stream.cuda_stream), C.byref(t)))
        return int(t.value)

    def upload_wait(self, ticket: int, compute_stream):
        capi.check(self.lib.lfe_uploader_stream_wait(self.host_uploader(), int(ticket), C.c_void_p(compute_stream.cuda_stream)))

    def upload_drain(self):
        if getattr(self, "_uploader", None) is not None:
            capi.check(self.lib.lfe_uploader_drain(self._uploader))

    def empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def zeros(self, n, dtype):
        return torch.zeros(int(n), dtype=dtype, device=self.device)

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _check_dev(self, *ts):
        for t in ts:
            if t is not None and (not t.is_cuda or not t.is_contiguous()):
                raise ValueError("expected contiguous CUDA tensors")

    def set_option(self, option: int, value: int):
        """lfe_set_option (capi.OPT_*)."""
        capi.check(self.lib.lfe_set_option(self.h, int(option), int(value)))

    # -- profiling -----------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        capi.check(self.lib.lfe_set_profiling(self.h, 1 if on else 0))

    def reset_timers(self):
        capi.check(self.lib.lfe_reset_timers(self.h))

    def timers(self) -> dict:
        out = (C.c_double * capi.N_TIMERS)()
        capi.check(self.lib.lfe_get_timers(self.h, out))
        return {capi.TIMER_NAMES[i]: out[i] for i in range(15)}

    def counters(self) -> dict:
        out = (C.c_int64 * 4)()
        capi.check(self.lib.lfe_get_counters(self.h, out))
        return {"launches": out[0], "sweep_launches": out[1], "collectives": out[2]}

    # -- (1) group codes ----------------------------------------------------------------------
    def factorize(self, keys: torch.Tensor, order: int = capi.ORDER_SORTED, want_uniques: bool = False,
                  uniques_cap: int | None = None):
        """-> (codes int32, n_groups, uniques or None)"""
        self._check_dev(keys)
        if keys.dtype not in _TORCH_TO_LFE:
            raise TypeError(f"unsupported key dtype {keys.dtype}")
        n = keys.numel()
        codes = self.empty(n, torch.int32)
        G = C.c_int64(0)
        uniq = None
        cap = 0
        if want_uniques:
            cap = n if uniques_cap is None else min(n, int(uniques_cap))
            if order & capi.ORDER_GLOBAL and self.world > 1:
                cap = int(uniques_cap or (1 << 20))     # the shared code space may hold keys this rank never sees
            cap = max(cap, 1)
            uniq = self.empty(cap, keys.dtype)
        rc = self.lib.lfe_factorize(self.h, self._p(keys), _TORCH_TO_LFE[keys.dtype], n, order, self._p(codes),
                                    C.byref(G), self._p(uniq), cap, self.stream)
        if rc == capi.ERR_CAPACITY and want_uniques and cap < int(G.value):
            # n_groups_out is filled even when the uniques did not fit: retry once with exactly that capacity (in the
            # shared code space of a multi-GPU fit G may exceed this rank's row count)
            return self.factorize(keys, order, True, int(G.value))
        capi.check(rc)
        return codes, int(G.value), (uniq[: G.value] if uniq is not None else None)

    def factorize_counts(self, codes: torch.Tensor, n_groups: int) -> torch.Tensor | None:
        """Group sizes left behind by the ``factorize`` call that produced ``codes`` (None when that call did not
        take the counting path); must be asked for right after it."""
        if n_groups <= 0:
            return None
        out = self.empty(n_groups, torch.int32)
        ok = C.c_int(0)
        capi.check(self.lib.lfe_factorize_counts(self.h, self._p(codes), n_groups, self._p(out), C.byref(ok), self.stream))
        return out if ok.value else None

    def group_counts(self, codes: torch.Tensor, n_groups: int) -> torch.Tensor:
        self._check_dev(codes)
        out = self.empty(n_groups, torch.int32)
        capi.check(self.lib.lfe_group_counts(self.h, self._p(codes), codes.numel(), n_groups, self._p(out), self.stream))
        return out

    def singleton_mask(self, codes: list[torch.Tensor], n_groups: list[int], local_fe: int = -1, counts=None):
        """-> (keep uint8, n_keep on this rank, fe_dims).  ``self.last_n_keep_global`` = kept rows over all ranks."""
        self._check_dev(*codes)
        n = codes[0].numel()
        keep = self.empty(n, torch.uint8)
        n_keep = (C.c_int64 * 3)()
        dims = (C.c_int64 * len(codes))()
        capi.check(self.lib.lfe_singleton_mask(self.h, capi.ptr_array([c.data_ptr() for c in codes]),
                                               capi.i64_array(n_groups), len(codes), n, int(local_fe),
                                               capi.ptr_array([c.data_ptr() for c in counts]) if counts else None,
                                               self._p(keep), n_keep, dims, self.stream))
        self.last_n_keep_global = int(n_keep[1])
        self.last_n_rows_global = int(n_keep[2])
        return keep, int(n_keep[0]), tuple(int(d) for d in dims)

    def compact(self, keep: torch.Tensor, n_keep: int, f64: list[torch.Tensor], i32: list[torch.Tensor]):
        self._check_dev(keep, *f64, *i32)
        f_out = [self.empty(n_keep, torch.float64) for _ in f64]
        i_out = [self.empty(n_keep, torch.int32) for _ in i32]
        capi.check(self.lib.lfe_compact(self.h, self._p(keep), keep.numel(),
                                        capi.ptr_array([t.data_ptr() for t in f64]),
                                        capi.ptr_array([t.data_ptr() for t in f_out]), len(f64),
                                        capi.ptr_array([t.data_ptr() for t in i32]),
                                        capi.ptr_array([t.data_ptr() for t in i_out]), len(i32), self.stream))
        return f_out, i_out

    def sort_rows(self, codes: torch.Tensor, n_groups: int):
        """-> (perm int32, already_sorted)"""
        self._check_dev(codes)
        perm = self.empty(codes.numel(), torch.int32)
        flag = C.c_int(0)
        capi.check(self.lib.lfe_sort_rows(self.h, self._p(codes), codes.numel(), n_groups, self._p(perm),
                                          C.byref(flag), self.stream))
        return perm, bool(flag.value)

    def gather(self, perm: torch.Tensor, f64: list[torch.Tensor], i32: list[torch.Tensor]):
        self._check_dev(perm, *f64, *i32)
        n = perm.numel()
        f_out = [self.empty(n, torch.float64) for _ in f64]
        i_out = [self.empty(n, torch.int32) for _ in i32]
        capi.check(self.lib.lfe_gather(self.h, self._p(perm), n, capi.ptr_array([t.data_ptr() for t in f64]),
                                       capi.ptr_array([t.data_ptr() for t in f_out]), len(f64),
                                       capi.ptr_array([t.data_ptr() for t in i32]),
                                       capi.ptr_array([t.data_ptr() for t in i_out]), len(i32), self.stream))
        return f_out, i_out

    # -- (2) demean ---------------------------------------------------------------------------
    def demean(self, cols: list[torch.Tensor], codes: list[torch.Tensor], n_groups: list[int], fe_order: list[int],
               weights: torch.Tensor | None = None, tol: float = 1e-6, max_iter: int = 50, min_iter: int = 3,
               local_fe: int = -1, strict_check: bool = False):
        """In place.  -> (iterations, max_mean, converged)"""
        self._check_dev(*cols, *codes, weights)
        n = cols[0].numel()
        for t in cols:
            if t.dtype != torch.float64 or t.numel() != n:
                raise ValueError("demean columns must be float64 of equal length")
        for t in codes:
            if t.dtype != torch.int32 or t.numel() != n:
                raise ValueError("codes must be int32 of the same length as the columns")
        it, mm, cv = C.c_int(0), C.c_double(0.0), C.c_int(0)
        capi.check(self.lib.lfe_demean(self.h, capi.ptr_array([t.data_ptr() for t in cols]), len(cols),
                                       capi.ptr_array([t.data_ptr() for t in codes]), capi.i64_array(n_groups),
                                       capi.i32_array(fe_order), len(codes), n, self._p(weights), float(tol),
                                       int(max_iter), int(min_iter), int(local_fe), 1 if strict_check else 0,
                                       C.byref(it), C.byref(mm),
                                       C.byref(cv), self.stream))
        return int(it.value), float(mm.value), bool(cv.value)

    def demean_plan(self, codes: list[torch.Tensor], n_groups: list[int], weights: torch.Tensor | None = None,
                    local_fe: int = -1, counts=None, columns_hint: int | None = None) -> "DemeanPlan":
        """Column-independent part of lfe_demean (entry lists, group sizes); see ``DemeanPlan.run``.
        ``counts`` (optional) = per-FE int32 group sizes from ``group_counts`` (saves one pass per FE);
        ``columns_hint`` = columns the plan will be run with at once (lets the library block the entry lists for
        the L2 when the biggest FE table does not fit it)."""
        return DemeanPlan(self, codes, n_groups, weights, local_fe, counts, columns_hint)

    # -- (4) gram / solve ---------------------------------------------------------------------
    def gram(self, cols: list[torch.Tensor], weights: torch.Tensor | None = None, add_intercept: bool = True):
        self._check_dev(*cols, weights)
        P = len(cols) + (1 if add_intercept else 0)
        n = cols[0].numel() if cols else 0
        out = self.empty(P * P, torch.float64)
        capi.check(self.lib.lfe_gram(self.h, capi.ptr_array([t.data_ptr() for t in cols]), len(cols), n,
                                     self._p(weights), 1 if add_intercept else 0, self._p(out), self.stream))
        return out.view(P, P)

    def solve_spd(self, A: torch.Tensor, b: torch.Tensor):
        """-> (beta, Ainv, used_fallback)"""
        A = A.contiguous()
        b = b.contiguous()
        self._check_dev(A, b)
        p = b.numel()
        beta = self.empty(p, torch.float64)
        Ainv = self.empty(p * p, torch.float64)
        fb = C.c_int(0)
        capi.check(self.lib.lfe_solve_spd(self.h, self._p(A), self._p(b), p, self._p(beta), self._p(Ainv),
                                          C.byref(fb), self.stream))
        return beta, Ainv.view(p, p), bool(fb.value)

    # -- (5) residuals + vcov ----------------------------------------------------------------
    def resid_vcov(self, cols: list[torch.Tensor], beta_full: torch.Tensor, Binv: torch.Tensor, vcov_kind: int,
                   cl_codes: list[torch.Tensor] | None, n_cl_groups: list[int] | None,
                   weights: torch.Tensor | None, ssc: bool, n_obs: int, df_resid: int, want_resid: bool = False):
        """-> dict(se, vcov, stats=[rss, tss, sum_w_r2], n_clusters, resid)"""
        cl_codes = cl_codes or []
        n_cl_groups = n_cl_groups or []
        self._check_dev(*cols, beta_full, weights, *cl_codes)
        k = len(cols) - 1
        n = cols[0].numel()
        Binv = Binv.contiguous() if Binv is not None else None
        se = self.empty(max(k, 1), torch.float64)
        V = self.empty(max(k * k, 1), torch.float64)
        stats = self.empty(4, torch.float64)
        resid = self.empty(n, torch.float64) if want_resid else None
        ncl = (C.c_int64 * max(len(cl_codes), 1))()
        capi.check(self.lib.lfe_resid_vcov(self.h, capi.ptr_array([t.data_ptr() for t in cols]), k, n,
                                           self._p(beta_full), self._p(Binv), vcov_kind,
                                           capi.ptr_array([t.data_ptr() for t in cl_codes]),
                                           capi.i64_array(n_cl_groups), len(cl_codes), self._p(weights),
                                           1 if ssc else 0, int(n_obs), int(df_resid), self._p(se), self._p(V),
                                           self._p(stats), self._p(resid), ncl, self.stream))
        return dict(se=se[:k], vcov=V[: k * k].view(k, k), stats=stats,
                    n_clusters=[int(ncl[i]) for i in range(len(cl_codes))], resid=resid)

    def yoco_vcov(self, cols: list[torch.Tensor], beta: torch.Tensor, Binv: torch.Tensor, nn, sum_y, sum_y_sq,
                  vcov_kind: int, cl_codes, n_cl_groups, ssc: bool, n_obs: int, df_resid: int):
        """-> dict(se, vcov, rss, n_clusters)"""
        cl_codes = cl_codes or []
        n_cl_groups = n_cl_groups or []
        Binv = Binv.contiguous()
        self._check_dev(*cols, beta, Binv, nn, sum_y, sum_y_sq, *cl_codes)
        k = len(cols) - 1
        se = self.empty(k, torch.float64)
        V = self.empty(k * k, torch.float64)
        rss = self.zeros(1, torch.float64)
        ncl = (C.c_int64 * max(len(cl_codes), 1))()
        capi.check(self.lib.lfe_yoco_vcov(self.h, capi.ptr_array([t.data_ptr() for t in cols]), k, cols[0].numel(),
                                          self._p(beta), self._p(Binv), self._p(nn), self._p(sum_y),
                                          self._p(sum_y_sq), vcov_kind,
                                          capi.ptr_array([t.data_ptr() for t in cl_codes]),
                                          capi.i64_array(n_cl_groups), len(cl_codes), 1 if ssc else 0, int(n_obs),
                                          int(df_resid), self._p(se), self._p(V), self._p(rss), ncl, self.stream))
        return dict(se=se, vcov=V.view(k, k), rss=rss, n_clusters=[int(ncl[i]) for i in range(len(cl_codes))])

    # -- (3) compress -------------------------------------------------------------------------
    def compress(self, key_codes: list[torch.Tensor], n_groups: list[int], y: torch.Tensor | None,
                 weights: torch.Tensor | None = None, out_cap: int | None = None, count_only: bool = False):
        """-> dict(n_compressed, key_codes, n, cnt, sum_y, sum_y_sq).  Key columns are int32 codes, or uint8
        raw values used as codes (n_groups = max + 1)."""
        self._check_dev(*key_codes, y, weights)
        n = key_codes[0].numel() if key_codes else (y.numel() if y is not None else 0)
        K = len(key_codes)
        for t in key_codes:
            if t.dtype not in (torch.int32, torch.uint8) or t.numel() != n or not t.is_contiguous():
                raise TypeError("compress: key columns must be contiguous int32 codes or uint8 values of equal length")
        dts = capi.i32_array([capi.LFE_U8 if t.dtype == torch.uint8 else capi.LFE_I32 for t in key_codes])
        G = C.c_int64(0)
        if count_only:
            capi.check(self.lib.lfe_compress(self.h, capi.ptr_array([t.data_ptr() for t in key_codes]), dts,
                                             capi.i64_array(n_groups), K, None, None, n, 0, None, None, None, None,
                                             None, C.byref(G), self.stream))
            return dict(n_compressed=int(G.value))
        space = 1
        for g in n_groups:
            space *= int(g)
        cap = int(min(n, space if out_cap is None else out_cap)) if K else 1
        while True:
            kout = [self.empty(cap, torch.int32) for _ in range(K)]
            nn, sy, syy = (self.empty(cap, torch.float64) for _ in range(3))
            cnt = self.empty(cap, torch.int32)
            rc = self.lib.lfe_compress(self.h, capi.ptr_array([t.data_ptr() for t in key_codes]), dts,
                                       capi.i64_array(n_groups), K, self._p(y), self._p(weights), n, cap,
                                       capi.ptr_array([t.data_ptr() for t in kout]), self._p(nn), self._p(cnt),
                                       self._p(sy), self._p(syy), C.byref(G), self.stream)
            if rc == capi.ERR_CAPACITY and cap < n:
                cap = n
                continue
            capi.check(rc)
            break
        g = int(G.value)
        return dict(n_compressed=g, key_codes=[t[:g] for t in kout], n=nn[:g], cnt=cnt[:g], sum_y=sy[:g],
                    sum_y_sq=syy[:g])

    # -- multi-GPU ---------------------------------------------------------------------------
    def comm_init_from_torch(self):
        """Attach an NCCL communicator spanning the default torch.distributed group."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        if world == 1:
            return
        uid = (C.c_ubyte * 128)()
        if rank == 0:
            capi.check(self.lib.lfe_comm_unique_id(uid))
        t = torch.tensor(list(uid), dtype=torch.uint8, device=self.device if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(t, src=0)
        buf = (C.c_ubyte * 128)(*t.cpu().tolist())
        capi.check(self.lib.lfe_comm_init(self.h, buf, rank, world))
        self.world, self.rank = world, rank

    def allreduce(self, t: torch.Tensor):
        capi.check(self.lib.lfe_allreduce_f64(self.h, self._p(t), t.numel(), self.stream))
        return t

    def expand_dummies(self, codes: torch.Tensor, levels: list[int], base: torch.Tensor | None = None,
                       as_uint8: bool = False) -> list[torch.Tensor]:
        """lfe_expand_dummies: the dummy columns (codes == level) [* base] of all ``levels`` in one pass."""
        self._check_dev(codes, base)
        n, J = codes.numel(), len(levels)
        if codes.dtype != torch.int32 or (base is not None and (base.dtype != torch.float64 or base.numel() != n)):
            raise ValueError("expand_dummies: int32 codes and a float64 base of the same length")
        dt = torch.uint8 if (as_uint8 and base is None) else torch.float64
        out = [self.empty(n, dt) for _ in range(J)]
        if J:
            capi.check(self.lib.lfe_expand_dummies(self.h, self._p(codes), n, capi.i32_array(levels), J, self._p(base),
                                                   _TORCH_TO_LFE[dt], capi.ptr_array([t.data_ptr() for t in out]),
                                                   self.stream))
        return out

    def shard_plan(self, codes: torch.Tensor, n_groups: int):
        """lfe_shard_plan: balanced code-range cuts of the shard key + this rank's send order.
        -> dict(cuts, send_counts, recv_counts, order)"""
        self._check_dev(codes)
        n, W = codes.numel(), self.world
        cuts, sc, rc = (C.c_int64 * (W + 1))(), (C.c_int64 * W)(), (C.c_int64 * W)()
        order = self.empty(n, torch.int32)
        capi.check(self.lib.lfe_shard_plan(self.h, self._p(codes), n, int(n_groups), cuts, sc, rc, self._p(order),
                                           self.stream))
        return dict(cuts=[int(c) for c in cuts], send_counts=[int(c) for c in sc], recv_counts=[int(c) for c in rc],
                    order=order)

    def shard_exchange(self, plan: dict, cols: list[torch.Tensor], i32_sub: list[int] | None = None):
        """lfe_shard_exchange: all-to-all of the columns by the plan's destination ranks.  -> new tensors holding
        the rows this rank owns (ordered by source rank, original order inside each source)."""
        self._check_dev(*cols)
        n = plan["order"].numel()
        n_recv = sum(plan["recv_counts"])
        for t in cols:
            if t.numel() != n or t.dtype not in _TORCH_TO_LFE:
                raise ValueError("shard_exchange: columns must have one value per planned row (int32/int64/uint8/float64)")
        out = [self.empty(n_recv, t.dtype) for t in cols]
        capi.check(self.lib.lfe_shard_exchange(
            self.h, self._p(plan["order"]), capi.i64_array(plan["send_counts"]), capi.i64_array(plan["recv_counts"]),
            capi.ptr_array([t.data_ptr() for t in cols]), capi.i32_array([_TORCH_TO_LFE[t.dtype] for t in cols]),
            capi.i64_array(i32_sub) if i32_sub is not None else None, len(cols), n,
            capi.ptr_array([t.data_ptr() for t in out]), self.stream))
        return out

    def key_ranges(self, cols: list[torch.Tensor]):
        """lfe_key_ranges: -> [(min, max)] of integer columns (uint8 / int32 / int64), one pass each, one sync;
        with a communicator attached, of the whole sharded columns."""
        self._check_dev(*cols)
        K, n = len(cols), cols[0].numel()
        lo, hi = (C.c_int64 * K)(), (C.c_int64 * K)()
        capi.check(self.lib.lfe_key_ranges(self.h, capi.ptr_array([t.data_ptr() for t in cols]),
                                           capi.i32_array([_TORCH_TO_LFE[t.dtype] for t in cols]), K, n, lo, hi,
                                           self.stream))
        return [(int(a), int(b)) for a, b in zip(lo, hi)]

    def compress_merge(self, part: dict, n_groups: list[int]):
        """lfe_compress_merge: all-gather + re-aggregate the ranks' partial YOCO tables (the dicts returned by
        :meth:`compress`); every rank gets the identical global table.  -> dict like :meth:`compress`"""
        keys = part["key_codes"]
        K, rows = len(keys), int(part["n_compressed"])
        self._check_dev(*keys, part["n"], part["cnt"], part["sum_y"], part["sum_y_sq"])
        space = 1
        for g in n_groups:
            space *= int(g)
        tot = self.empty(1, torch.float64)
        tot.fill_(float(rows))
        cap = int(min(space, int(self.allreduce(tot).item()))) if self.world > 1 else rows
        cap = max(cap, 1)
        kout = [self.empty(cap, torch.int32) for _ in range(K)]
        nn, sy, syy = (self.empty(cap, torch.float64) for _ in range(3))
        cnt = self.empty(cap, torch.int32)
        G = C.c_int64(0)
        capi.check(self.lib.lfe_compress_merge(
            self.h, capi.ptr_array([t.data_ptr() for t in keys]), capi.i64_array(n_groups), K, self._p(part["n"]),
            self._p(part["cnt"]), self._p(part["sum_y"]), self._p(part["sum_y_sq"]), rows, cap,
            capi.ptr_array([t.data_ptr() for t in kout]), self._p(nn), self._p(cnt), self._p(sy), self._p(syy),
            C.byref(G), self.stream))
        g = int(G.value)
        return dict(n_compressed=g, key_codes=[t[:g] for t in kout], n=nn[:g], cnt=cnt[:g], sum_y=sy[:g],
                    sum_y_sq=syy[:g])

    def replicated(self):
        """Context manager: the calls inside run rank-locally although a communicator is attached (work that
        every rank repeats on identical data, e.g. the solve on the merged YOCO table)."""
        return _Replicated(self)


class _Replicated:
    def __init__(self, eng):
        self.eng = eng

    def __enter__(self):
        capi.check(self.eng.lib.lfe_comm_set_active(self.eng.h, 0))
        self.saved, self.eng.world = self.eng.world, 1
        return self.eng

    def __exit__(self, *exc):
        self.eng.world = self.saved
        capi.check(self.eng.lib.lfe_comm_set_active(self.eng.h, 1))
        return False


class DemeanPlan:
    """lfe_demean_plan_*: built from the code columns only, runnable on any number of column sets."""

    def __init__(self, eng: Engine, codes, n_groups, weights, local_fe, counts=None, columns_hint=None):
        eng._check_dev(*codes, weights)
        self.eng, self.codes, self.weights, self.n_groups = eng, list(codes), weights, list(n_groups)   # keep alive
        self.n = codes[0].numel()
        for t in codes:
            if t.dtype != torch.int32 or t.numel() != self.n:
                raise ValueError("codes must be int32 tensors of equal length")
        p = C.c_void_p()
        eng.set_option(capi.OPT_COLUMNS_HINT, int(columns_hint or 0))
        capi.check(eng.lib.lfe_demean_plan_create(eng.h, capi.ptr_array([t.data_ptr() for t in codes]),
                                                  capi.i64_array(n_groups), len(codes), self.n, eng._p(weights),
                                                  int(local_fe),
                                                  capi.ptr_array([c.data_ptr() for c in counts]) if counts else None,
                                                  C.byref(p), eng.stream))
        self.p = p

    def run(self, cols: list[torch.Tensor], fe_order: list[int], tol: float = 1e-6, max_iter: int = 50,
            min_iter: int = 3, strict_check: bool = False, src: list[torch.Tensor] | None = None):
        """Demean ``cols`` (cols[0] = y) in place, or -- with ``src`` -- read the inputs from ``src`` and write the
        demeaned columns to ``cols`` (caller data untouched, no extra copy).  -> (iterations, max_mean, converged)"""
        eng = self.eng
        eng._check_dev(*cols, *(src or []))
        if src is not None and (len(src) != len(cols) or any(t.dtype != torch.float64 or t.numel() != self.n for t in src)):
            raise ValueError("src must match cols (float64, one value per plan row)")
        for t in cols:
            if t.dtype != torch.float64 or t.numel() != self.n:
                raise ValueError("demean columns must be float64 with one value per plan row")
        it, mm, cv = C.c_int(0), C.c_double(0.0), C.c_int(0)
        capi.check(eng.lib.lfe_demean_plan_run(eng.h, self.p, capi.ptr_array([t.data_ptr() for t in cols]),
                                               capi.ptr_array([t.data_ptr() for t in src]) if src is not None else None,
                                               len(cols),
                                               capi.i32_array(fe_order), float(tol), int(max_iter), int(min_iter),
                                               1 if strict_check else 0, C.byref(it), C.byref(mm), C.byref(cv),
                                               eng.stream))
        return int(it.value), float(mm.value), bool(cv.value)

    FIT_MAX_COLS = 11    # TL_MAXC of csrc/tail.cuh: y + regressors handled by lfe_demean_plan_fit

    def fit(self, cols: list[torch.Tensor], fe_order: list[int], tol: float, max_iter: int, min_iter: int,
            vcov_kind: int, cl_codes: torch.Tensor | None, n_cl_groups: int, ssc: bool, n_obs: int, df_resid: int):
        """lfe_demean_plan_fit: sweeps + normal equations + solve + residual statistics + IID / HC1 / one-way
        clustered SEs without materialising the demeaned columns (``cols`` are only read; cols[0] = y).
        -> dict(beta_full, se, vcov, stats, iterations, max_mean, converged, n_clusters)"""
        eng = self.eng
        eng._check_dev(*cols, cl_codes)
        k = len(cols) - 1
        for t in cols:
            if t.dtype != torch.float64 or t.numel() != self.n:
                raise ValueError("fit columns must be float64 with one value per plan row")
        if cl_codes is not None and (cl_codes.dtype != torch.int32 or cl_codes.numel() != self.n):
            raise ValueError("cluster codes must be int32 with one value per plan row")
        out = eng.empty((k + 1) + max(k, 1) + max(k * k, 1) + 4, torch.float64)
        beta, se = out[: k + 1], out[k + 1: k + 1 + max(k, 1)]
        V = out[k + 1 + max(k, 1): k + 1 + max(k, 1) + max(k * k, 1)]
        stats = out[-4:]
        it, mm, cv, ncl = C.c_int(0), C.c_double(0.0), C.c_int(0), C.c_int64(0)
        capi.check(eng.lib.lfe_demean_plan_fit(
            eng.h, self.p, capi.ptr_array([t.data_ptr() for t in cols]), len(cols), capi.i32_array(fe_order), float(tol),
            int(max_iter), int(min_iter), 0, int(vcov_kind), eng._p(cl_codes), int(n_cl_groups), 1 if ssc else 0,
            int(n_obs), int(df_resid), eng._p(beta), eng._p(se), eng._p(V), eng._p(stats), C.byref(it), C.byref(mm),
            C.byref(cv), C.byref(ncl), eng.stream))
        return dict(beta_full=beta, se=se[:k], vcov=V[: k * k].view(k, k), stats=stats, packed=out,
                    iterations=int(it.value), max_mean=float(mm.value), converged=bool(cv.value), n_clusters=int(ncl.value))

    def close(self):
        if getattr(self, "p", None):
            self.eng.lib.lfe_demean_plan_destroy(self.eng.h, self.p)
            self.p = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_ENGINES: dict[int, Engine] = {}


def get_engine(device: int | None = None) -> Engine:
    _require_cuda()
    idx = torch.cuda.current_device() if device is None else int(device)
    if idx not in _ENGINES:
        _ENGINES[idx] = Engine(idx)
    return _ENGINES[idx]


def to_numpy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()
