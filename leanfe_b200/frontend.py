"""Input adapters and column expansion (the reference's front end, polars_impl.py:311-365).

Host-side, metadata-level work stays in Python: resolving columns from a dict / pandas / pyarrow /
parquet / Polars frame, dictionary-encoding string ids (the reference's ``_cats_to_int``,
polars_impl.py:118-139, is also a host-side hash), naming dummy columns.  Everything per-row runs on
the device.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _capi as capi
from .engine import Engine


class Frame:
    """Uniform, lazy column access over the accepted ``data`` kinds."""

    def __init__(self, data, needed: list[str]):
        self._cols = {}
        self._getter = None
        if data is None:
            raise ValueError("A dataset must be provided")
        if isinstance(data, str):
            import pyarrow.parquet as pq
            table = pq.read_table(data, columns=needed)
            self._from_arrow(table)
        elif isinstance(data, dict):
            self._getter = data.__getitem__
            self._names = list(data.keys())
        else:
            mod = type(data).__module__.split(".")[0]
            if mod == "pandas":
                self._getter = lambda c: data[c].to_numpy()
                self._names = list(data.columns)
            elif mod == "pyarrow":
                self._from_arrow(data.select([c for c in needed if c in data.column_names]))
            elif mod == "polars":
                lf = data.lazy() if hasattr(data, "lazy") else data
                self._from_arrow(lf.select(needed).collect().to_arrow())
            else:
                raise TypeError(f"unsupported data type {type(data)!r}: pass a dict of arrays/tensors, a pandas "
                                "DataFrame, a pyarrow Table, a Polars frame or a parquet path")
        missing = [c for c in needed if c not in self._names]
        if missing:
            raise KeyError(f"columns not found in data: {missing}")

    def _from_arrow(self, table):
        self._names = list(table.column_names)
        self._getter = lambda c: _arrow_to_numpy(table.column(c))

    def column(self, name: str):
        if name not in self._cols:
            self._cols[name] = self._getter(name)
        return self._cols[name]


def _arrow_to_numpy(chunked):
    import pyarrow as pa
    arr = chunked.combine_chunks() if hasattr(chunked, "combine_chunks") else chunked
    if arr.null_count:
        raise ValueError("null values are not supported in model columns (the reference's Polars path never filters them)")
    if pa.types.is_dictionary(arr.type):
        arr = arr.dictionary_decode()
    if pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type):
        return np.asarray(arr.to_pylist(), dtype=object)
    return arr.to_numpy(zero_copy_only=False)


_NP_TO_TORCH = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32, np.dtype(np.int64): torch.int64,
                np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8}


class Uploader:
    """All host->device copies of a fit, enqueued up front, key columns first, so that factorisation, the singleton
    test and the entry-list sorts (which need only the ids) run while the fp64 columns are still crossing PCIe; the
    compute stream waits on a per-column event right before a column's first use.

    Pinned host tensors are copied with one asynchronous ``copy_`` each on the engine's copy stream.  PAGEABLE
    sources -- NumPy arrays, pandas / Arrow columns, unpinned tensors: what the reference API receives
    (polars_impl.py:342-347) -- go through the library's pinned ring (``lfe_uploader_*``, csrc/upload.cu): host
    threads stage 32 MB chunks, one cudaMemcpyAsync per chunk; read-only sources are read where they lie.
    """

    def __init__(self, eng: Engine, frame: Frame, names: list[str]):
        self.eng, self.frame = eng, frame
        self.compute = torch.cuda.current_stream(eng.device)
        self.copy = None
        self.dev: dict[str, torch.Tensor] = {}
        self.events: dict[str, object] = {}    # name -> torch Event (pinned path) or int ticket (pinned-ring path)
        self._host_refs: list = []
        self._ring_used = False
        self.events_seen: set[str] = set()     # columns whose device copy was made here (never caller-owned)
        jobs = []
        for name in dict.fromkeys(names):
            col = frame.column(name)
            if isinstance(col, torch.Tensor):
                if col.is_cuda:
                    self.dev[name] = col.contiguous()
                    continue
                t = col.contiguous()
                jobs.append((name, t, t.dtype, tuple(t.shape), t.data_ptr() if not t.is_pinned() else None))
            elif _is_stringlike(col):
                continue                                   # dictionary-encoded on the host at first use
            else:
                a = np.asarray(col)
                if a.dtype == np.bool_:
                    a = a.astype(np.uint8)
                elif a.dtype in (np.int8, np.int16, np.uint16):
                    a = a.astype(np.int32)
                elif a.dtype == np.float16:
                    a = a.astype(np.float32)
                a = np.ascontiguousarray(a)                # no-op for the usual contiguous column
                if a.dtype in _NP_TO_TORCH:
                    jobs.append((name, a, _NP_TO_TORCH[a.dtype], a.shape, a.ctypes.data))
                else:                                      # exotic dtype: let torch convert (needs a writable array)
                    t = torch.from_numpy(a if a.flags.writeable else a.copy())
                    jobs.append((name, t, t.dtype, tuple(t.shape), t.data_ptr()))
        if jobs:
            self.copy = eng.copy_stream()
            self.copy.wait_stream(self.compute)
            for name, src, dtype, shape, ptr in jobs:
                # destination allocated on the COMPUTE stream (so the caching allocator reuses the previous fit's
                # blocks instead of growing a per-copy-stream pool), filled on the engine's one copy stream
                d = torch.empty(shape, dtype=dtype, device=eng.device)
                d.record_stream(self.copy)
                if ptr is None:                            # pinned tensor: one truly asynchronous copy
                    with torch.cuda.stream(self.copy):
                        d.copy_(src, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self.copy)
                    self.events[name] = ev
                else:                                      # pageable: the library's pinned ring
                    self.events[name] = eng.upload_enqueue(ptr, d.numel() * d.element_size(), d, self.copy)
                    self._ring_used = True
                self.dev[name] = d
                self.events_seen.add(name)
                self._host_refs.append(src)                # keep the source alive until its copy ran

    def get(self, name: str):
        """Device tensor of an uploaded column (compute stream ordered after its copy), else the host column."""
        if name in self.dev:
            ev = self.events.pop(name, None)
            if isinstance(ev, int):
                self.eng.upload_wait(ev, self.compute)
            elif ev is not None:
                self.compute.wait_event(ev)
            return self.dev[name]
        return self.frame.column(name)

    def close(self):
        """End of the fit: every ring job has completed before the host sources may go away."""
        if self._ring_used:
            self.eng.upload_drain()
            self._ring_used = False
        self._host_refs.clear()


@dataclass
class KeyColumn:
    """A group-key column on the device: dense int32 codes + cardinality (+ labels when needed)."""
    codes: torch.Tensor
    n_groups: int
    labels: list | None = None     # python scalars in code order (sorted ascending)
    counts: torch.Tensor | None = None   # int32 group sizes when the factorisation produced them for free


def _is_stringlike(col) -> bool:
    return isinstance(col, np.ndarray) and col.dtype.kind in ("O", "U", "S")


def to_device_f64(eng: Engine, col) -> torch.Tensor:
    if isinstance(col, torch.Tensor):
        t = col.to(eng.device, non_blocking=True)
    else:
        a = np.asarray(col)
        if a.dtype.kind in ("O", "U", "S"):
            raise TypeError("string column used as a numeric regressor")
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(eng.device, non_blocking=True)
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t.contiguous()


def to_device_key(eng: Engine, col, want_labels: bool = False, global_codes: bool = False) -> KeyColumn:
    """Factorise a key column to sorted-rank codes (== np.unique(..., return_inverse=True)).

    ``global_codes`` (multi-GPU): codes come from one code space shared by all ranks (LFE_ORDER_GLOBAL).
    """
    if _is_stringlike(col):
        if global_codes:
            raise NotImplementedError("multi-GPU fits need numeric ids for the non-shard key columns")
        # strings stay on the host: dictionary-encode there, ship int32 codes (SURVEY 7 'hard parts')
        uniq, inv = np.unique(col, return_inverse=True)
        codes = torch.from_numpy(inv.reshape(-1).astype(np.int32)).to(eng.device, non_blocking=True)
        return KeyColumn(codes, int(uniq.shape[0]), uniq.tolist() if want_labels else None)
    if isinstance(col, torch.Tensor):
        t = col.to(eng.device, non_blocking=True).contiguous()
    else:
        a = np.asarray(col)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        elif a.dtype in (np.int8, np.int16):
            a = a.astype(np.int32)
        elif a.dtype == np.uint16:
            a = a.astype(np.int32)
        elif a.dtype == np.float32:
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(eng.device, non_blocking=True)
    if t.dtype in (torch.int8, torch.int16):
        t = t.to(torch.int32)
    elif t.dtype in (torch.float32, torch.float16, torch.bfloat16):
        t = t.to(torch.float64)
    elif t.dtype == torch.bool:
        t = t.to(torch.uint8)
    order = capi.ORDER_SORTED | (capi.ORDER_GLOBAL if global_codes else 0)
    codes, G, uniq = eng.factorize(t, order, want_uniques=want_labels, uniques_cap=1 << 16)
    counts = eng.factorize_counts(codes, G)
    labels = uniq.cpu().tolist() if want_labels else None
    return KeyColumn(codes, G, labels, counts)


def resolve_reference(labels: list, ref, var: str):
    """Reference level of ``i(var, ref=...)``: first sorted level by default (polars_impl.py:44-55)."""
    if ref is None:
        return labels[0]
    cand = ref
    if labels and not isinstance(labels[0], type(ref)):
        try:
            cand = type(labels[0])(ref)
        except (ValueError, TypeError):
            pass
    if cand not in labels:
        raise ValueError(f"Reference category '{ref}' not found in {var}. Available: {labels}")
    return cand
