"""Batched evaluation around the forward - what the reference's per-video Python loop does
(train/ucf_test.py:70-178, 336-353; train/xd_test.py; test.py:76-174), restructured for one or many B200s:

  * videos are chunked once with the reference's rule (data/tools.py:100-114: int(T/256)+1 zero-padded chunks of
    256, an extra all-zero chunk when T % 256 == 0) and all chunks of all videos of a rank go through ONE forward
    call `[sum S_v, 256, D]` (the library slabs it internally) instead of one launch sequence per video;
  * the valid rows (`logits[:len_cur]`, :112-114) are compacted on the device, never concatenated on the host;
  * multi-GPU: videos are independent, so ranks take disjoint video subsets (longest-processing-time greedy on the
    chunk count) with no data-path collective; the only exchange is one `all_gather` of the padded per-rank score
    vectors, after which every rank re-orders them into list order and computes AUC / AP, Ano-AUC and the class-wise
    AUC / AP with the radix-sort + scan kernels (exact-integer tie handling, float64 final division).

The class only orchestrates: chunking indices are built once on the host, every per-frame operation runs in
libiefvad.so kernels."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import os

import numpy as np
import torch

from . import _lib, ops

NORMAL_KEYS = ("Normal", "normal")


def num_chunks(T: int, maxlen: int = 256) -> int:
    """data/tools.py:100-114 + train/ucf_test.py:79-81."""
    return 1 if T < maxlen else int(T / maxlen) + 1


def partition_videos(lengths: Sequence[int], world: int, maxlen: int = 256) -> List[List[int]]:
    """Deterministic LPT assignment of videos to ranks by chunk count (ties by index)."""
    order = sorted(range(len(lengths)), key=lambda v: (-num_chunks(int(lengths[v]), maxlen), v))
    load = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for v in order:
        r = min(range(world), key=lambda i: (load[i], i))
        parts[r].append(v)
        load[r] += num_chunks(int(lengths[v]), maxlen)
    for p in parts:
        p.sort()
    return parts


_NVTX = os.environ.get("IEFVAD_NVTX", "0") not in ("", "0")


class _nvtx:
    """NVTX range around a phase of an evaluation pass (IEFVAD_NVTX=1; the library marks the forward's stages itself)."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def gather_layout(lengths: Sequence[int], parts: List[List[int]]):
    """Index arithmetic of the single collective.  Rank r packs the valid scores of its videos back to back into a
    vector padded to max_count; after all_gather the concatenation [world * max_count] is re-ordered into list order
    by segment copies (src offset in the gathered buffer, dst offset in list order, length) - one entry per video.
    Pure host function (numpy), shared by the CUDA evaluator and the CPU/gloo test of the sharding logic."""
    lengths = np.asarray(lengths, dtype=np.int64)
    global_off = np.concatenate([[0], np.cumsum(lengths)])[:-1]
    counts = [int(lengths[p].sum()) if len(p) else 0 for p in parts]
    max_count = max(counts) if counts else 0
    src, dst, ln = [], [], []
    for r, p in enumerate(parts):
        off = 0
        for v in p:
            src.append(r * max(max_count, 1) + off)
            dst.append(int(global_off[v]))
            ln.append(int(lengths[v]))
            off += int(lengths[v])
    return (np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64), np.asarray(ln, dtype=np.int64),
            counts, max_count)


def _segment_copy(src, src_off, dst, dst_off, length):
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib.iefvad_segment_copy(src.data_ptr(), src_off.data_ptr(), dst.data_ptr(), dst_off.data_ptr(),
                                                length.data_ptr(), src_off.numel(),
                                                torch.cuda.current_stream(src.device).cuda_stream))


class Evaluator:
    """Frame-level evaluation of one video list on `world` GPUs (one process per GPU).

    lengths[v] = embedding rows T_v, classes[v] = class key, gt = 0/1 per raw frame (16 per row) in list order."""

    def __init__(self, model, lengths: Sequence[int], classes: Sequence[str], gt, *, maxlen: int = 256, repeat: int = 16,
                 rank: int = 0, world: int = 1, device: Optional[torch.device] = None, process_group=None):
        self.model = model
        self.maxlen, self.repeat = maxlen, repeat
        self.rank, self.world, self.group = rank, world, process_group
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lengths = np.asarray(lengths, dtype=np.int64)
        self.classes = list(classes)
        n = len(self.lengths)
        assert len(self.classes) == n
        self.total_rows = int(self.lengths.sum())
        gt = np.asarray(gt)
        if gt.size != repeat * self.total_rows:
            raise ValueError(f"gt has {gt.size} frames, expected {repeat} x {self.total_rows}")
        self.global_off = np.concatenate([[0], np.cumsum(self.lengths)])[:-1]
        self.parts = partition_videos(self.lengths, world, maxlen)
        self.mine = self.parts[rank]
        g_src, g_dst, g_len, self.counts, self.max_count = gather_layout(self.lengths, self.parts)
        dev = self.device
        i64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int64), device=dev)  # noqa: E731
        # local layout: chunk-row offset of each of my videos inside my [sum S, maxlen] score matrix
        my_chunks = np.array([num_chunks(int(self.lengths[v]), maxlen) for v in self.mine], dtype=np.int64)
        self.local_chunks = int(my_chunks.sum())
        self.my_rows = int(self.lengths[self.mine].sum()) if len(self.mine) else 0
        self.chunk_row0 = np.concatenate([[0], np.cumsum(my_chunks)])[:-1] * maxlen
        self._src_off = i64(self.chunk_row0)
        self._dst_off = i64(np.concatenate([[0], np.cumsum(self.lengths[self.mine])])[:-1] if len(self.mine) else [])
        self._len = i64(self.lengths[self.mine] if len(self.mine) else [])
        # gather layout: rank r's padded vector holds its videos back to back
        self._g_src, self._g_dst, self._g_len = i64(g_src), i64(g_dst), i64(g_len)
        # labels: positives per embedding row (gt is 16 raw frames per row, list/ucf_generate_gt.py:24)
        self.pos = torch.as_tensor(gt.reshape(-1, repeat).sum(axis=1).astype(np.int32), device=dev)
        # class-wise / Ano-AUC subsets (train/ucf_test.py:164-178, 336-353) as one membership bit mask per row:
        # bit 0 = every row, bit 1 = rows of abnormal videos, bit 2 + c = rows of the c-th class key (list order)
        self.class_keys = list(dict.fromkeys(self.classes))
        if len(self.class_keys) > 30:
            raise ValueError("at most 30 class keys (32 subsets per ranking pass)")
        mask = np.zeros(self.total_rows, dtype=np.int64)
        for v in range(n):
            bits = 1 | (0 if self.classes[v] in NORMAL_KEYS else 2) | (4 << self.class_keys.index(self.classes[v]))
            mask[int(self.global_off[v]):int(self.global_off[v] + self.lengths[v])] = bits
        self.member = torch.as_tensor(mask.astype(np.uint32).view(np.int32), device=dev)
        self.num_subsets = 2 + len(self.class_keys)
        # valid-rows mode: per chunk the number of real rows (a prefix; 0 for the all-zero chunk of a T % maxlen == 0
        # video) and the row map b * maxlen + t of every real row - in video order, so the compact scores of the
        # forward ARE the packed per-rank score vector (no compaction pass)
        chunk_valid: List[int] = []
        for v in self.mine:
            T = int(self.lengths[v])
            for k in range(num_chunks(T, maxlen)):
                chunk_valid.append(max(0, min(maxlen, T - k * maxlen)))
        self._chunk_valid = chunk_valid
        rm = np.concatenate([np.arange(n_, dtype=np.int64) + c * maxlen for c, n_ in enumerate(chunk_valid)]
                            ) if chunk_valid else np.zeros(0, dtype=np.int64)
        self._rowmap = torch.as_tensor(rm.astype(np.int32), device=dev)
        cv = np.asarray(chunk_valid, dtype=np.int64)
        self._chunk_start = torch.as_tensor(np.concatenate([[0], np.cumsum(cv)])[:-1].astype(np.int64) if len(cv) else
                                            np.zeros(0, dtype=np.int64), device=dev)
        self._chunk_valid_dev = torch.as_tensor(cv.astype(np.int32), device=dev)
        self._ragged = None
        self.fell_back = False                 # a pass was redone under plan B after an fp16 overflow (on_overflow = "fallback")
        self.valid_rows_only = True
        self._img = self._ev = None
        self._pinned = None

    # ------------------------------------------------------------------ feature staging
    def chunk_features(self, feats: Sequence[torch.Tensor], dtype=torch.float16, pin: bool = False) -> torch.Tensor:
        """Host: my videos' [T_v, D] features -> one zero-padded [sum S_v, maxlen, D] tensor (process_split rule)."""
        D = feats[0].shape[1] if len(feats) else self.model.embed_dim
        out = torch.zeros((self.local_chunks, self.maxlen, D), dtype=dtype, pin_memory=pin)
        flat = out.view(-1, D)
        for i, f in enumerate(feats):
            r0 = int(self.chunk_row0[i])
            flat[r0:r0 + f.shape[0]] = torch.nan_to_num(f, nan=0.0).to(dtype)     # train/ucf_test.py:83-88
        return out

    def set_device_features(self, img_chunks: torch.Tensor, ev_chunks: torch.Tensor) -> None:
        self._img, self._ev = img_chunks.to(self.device), ev_chunks.to(self.device)

    def set_host_features(self, img_chunks: torch.Tensor, ev_chunks: torch.Tensor) -> None:
        """Pinned host staging for the end-to-end path (H2D inside every step)."""
        self._pinned = (img_chunks if img_chunks.is_pinned() else img_chunks.pin_memory(),
                        ev_chunks if ev_chunks.is_pinned() else ev_chunks.pin_memory())
        self._img = self._ev = None

    def set_host_ragged(self, feats_img: Sequence[torch.Tensor], feats_ev: Sequence[torch.Tensor],
                        dtype=torch.float16) -> None:
        """Pinned host staging of my videos' RAW [T_v, D] features back to back (what the .npy files hold): the
        end-to-end path then copies only real rows; the chunk / zero-pad rule is applied on the device (row N2)."""
        def pack(feats):
            rows = [torch.nan_to_num(f, nan=0.0).to(dtype) for f in feats]             # train/ucf_test.py:83-88
            out = torch.empty((self.my_rows, self.model.embed_dim), dtype=dtype, pin_memory=True)
            if rows:
                torch.cat(rows, out=out)
            return out
        if [int(f.shape[0]) for f in feats_img] != [int(self.lengths[v]) for v in self.mine]:
            raise ValueError("features do not match this rank's video lengths")
        self._ragged = (pack(feats_img), pack(feats_ev))
        self._pinned = None
        self._img = self._ev = None

    # ------------------------------------------------------------------ one evaluation pass
    def local_scores(self, host_inputs: bool = False, extra: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """Forward over all my chunks -> compacted sigmoid scores of my valid rows (device, fp32).  With host_inputs the
        pinned host features go through the library's pipelined host-input forward (copy overlapped with compute).
        extra: optional compact outputs of the valid-rows forward (see `step(extras=...)`)."""
        packed = torch.empty(max(self.max_count, 1), dtype=torch.float32, device=self.device)
        if extra and not (self.local_chunks and self.valid_rows_only and str(self.model.temporal.precision) != "fp32"):
            raise RuntimeError("extras need the valid-rows evaluation forward (a tensor-core precision plan)")
        # (the fp32 FFMA plan has no valid-rows mode: it runs the full forward and compacts afterwards)
        if self.local_chunks and self.valid_rows_only and str(self.model.temporal.precision) != "fp32":
            if host_inputs and self._ragged is not None:
                out = self.model.temporal.scores_ragged(self._ragged[0], self._ragged[1], self.device, self.maxlen,
                                                        self._chunk_valid, self._rowmap, self._chunk_start,
                                                        self._chunk_valid_dev, extra=extra)
            else:
                img, ev = self._pinned if host_inputs else (self._img, self._ev)
                out = self.model.temporal.scores(img, ev, self.device, self._chunk_valid, self._rowmap, extra=extra)
            self._keep = out["_keepalive"]
            packed[:self.my_rows].copy_(out["scores"])
            return packed
        if self.local_chunks:
            if host_inputs:
                out = self.model.temporal.scores_from_host(self._pinned[0], self._pinned[1], self.device)
            else:
                out = self.model.temporal(self._img, self._ev, with_scores=True)
            _segment_copy(out["scores"], self._src_off, packed, self._dst_off, self._len)
        return packed

    def gather(self, packed: torch.Tensor) -> torch.Tensor:
        """The single collective: all ranks' padded score vectors -> list-order score vector on every rank.  (Also used
        for the per-frame means of the fusion weights when `step(extras=...)` asks for them.)"""
        with _nvtx("iefvad gather"):
            return self._gather(packed)

    def _gather(self, packed: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            import torch.distributed as dist
            allv = torch.empty(self.world * max(self.max_count, 1), dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(allv, packed, group=self.group)
        else:
            allv = packed
        scores = torch.empty(self.total_rows, dtype=torch.float32, device=self.device)
        if self.total_rows:
            _segment_copy(allv, self._g_src, scores, self._g_dst, self._g_len)
        return scores

    def metrics_async(self, scores: torch.Tensor) -> torch.Tensor:
        """AUC / AP overall, Ano-AUC and class-wise AUC / AP from one ranking pass; the [num_subsets, 4] table is
        copied to pinned host memory on the stream - nothing here waits for the device."""
        with _nvtx("iefvad ranking"):
            return self._metrics_async(scores)

    def _metrics_async(self, scores: torch.Tensor) -> torch.Tensor:
        table = ops.auc_ap_multi(scores, self.pos, self.member, self.num_subsets, self.repeat)
        # a ring of pinned tables: a fresh pinned allocation per pass is a cudaHostAlloc whenever the previous ones are still
        # in flight (a loop of asynchronous passes), which serialises host and device
        ring = self.__dict__.setdefault("_table_ring", [])
        i = self.__dict__.get("_table_i", 0)
        self._table_i = i + 1
        if len(ring) < 64:
            ring.append([torch.empty(table.shape, dtype=table.dtype, pin_memory=True), torch.cuda.Event()])
            host, done = ring[-1]
        else:
            host, done = ring[i % 64]
            done.synchronize()                      # the copy of 64 passes ago (long finished unless the caller never syncs)
            if host.shape != table.shape:
                host = ring[i % 64][0] = torch.empty(table.shape, dtype=table.dtype, pin_memory=True)
        host.copy_(table, non_blocking=True)
        done.record(torch.cuda.current_stream(self.device))
        return host

    def finish(self, host_table: torch.Tensor) -> Dict[str, object]:
        """Wait for the stream and unpack a metrics table returned by `metrics_async`."""
        side = self.__dict__.get("_side_stream")
        if side is not None:
            side.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
        if not self.model.temporal.check_finite():          # range guard of the 16-bit plans (fp16 saturates at 65 504)
            self.model.temporal._raise_overflow()
        table = host_table.numpy()
        res: Dict[str, object] = {"AUC": float(table[0, 0]), "AP": float(table[0, 1]),
                                  "ano_AUC": float(table[1, 0]),      # NaN when one label value only (:347-350)
                                  "classwise": {}}
        for c, name in enumerate(self.class_keys):
            row = table[2 + c]
            if row[2] > 0:                              # classes without positives are skipped (:167-168)
                res["classwise"][name] = (float(row[0]), float(row[1]))
        return res

    def metrics(self, scores: torch.Tensor) -> Dict[str, object]:
        return self.finish(self.metrics_async(scores))

    def step(self, host_inputs: bool = False, with_metrics: bool = True, sync: bool = True,
             extras: Sequence[str] = (), overlap: bool = False) -> Dict[str, object]:
        """One evaluation pass.  With sync=False nothing waits for the device: the returned dict holds the device
        score vector and, under "pending", the pinned metrics table to hand to `finish()` later.

        extras - what the reference's loop collects per class besides the scores (train/ucf_test.py:124-144), kept on the
        device: "w_mean" adds "wi_mean" / "we_mean" ([total rows] in list order on every rank: w_i.mean(-1), w_e.mean(-1)
        of every frame, reduced inside the fusion kernel) and "classwise_wi" / "classwise_we" (class key -> the frames of
        that class's videos, list order); "wide" adds this rank's compact [my rows, D] "fused", "image_mu", "event_mu"
        (rows of `self.mine` videos back to back) and "classwise_fused" / "classwise_image_mu" / "classwise_event_mu" for
        the classes of this rank's videos.

        overlap (with sync=False): the collective and the ranking of this pass run on a side stream, so in a loop of passes
        they hide behind the NEXT pass's forward (they are a handful of latency-bound launches that would otherwise sit
        between two forwards on every rank); `finish()` waits for the side stream."""
        if overlap and not sync and with_metrics and not extras:
            return self._step_overlapped(host_inputs)
        extra: Dict[str, torch.Tensor] = {}
        unknown = set(extras) - {"w_mean", "wide"}
        if unknown:
            raise KeyError(f"unknown extras {sorted(unknown)}; choose from 'w_mean', 'wide'")
        if "w_mean" in extras:
            for k in ("wi_mean", "we_mean"):
                extra[k] = torch.zeros(max(self.max_count, 1), dtype=torch.float32, device=self.device)
        if "wide" in extras:
            for k in ("fused", "image_mu", "event_mu"):
                extra[k] = torch.empty((self.my_rows, self.model.embed_dim), dtype=torch.float32, device=self.device)
        call_extra = {k: (v[:self.my_rows] if v.dim() == 1 else v) for k, v in extra.items()}
        scores = self.gather(self.local_scores(host_inputs, call_extra or None))
        temporal = self.model.temporal
        if sync and getattr(temporal, "on_overflow", "raise") == "fallback" and str(temporal.precision) not in ("B", "fp32"):
            if not temporal.check_finite():                  # an fp16 operand overflowed: redo the pass in bf16 (plan B)
                keep = temporal.precision
                temporal.precision = "B"
                try:
                    scores = self.gather(self.local_scores(host_inputs))
                finally:
                    temporal.precision = keep
                self.fell_back = True
        res: Dict[str, object] = {}
        if with_metrics:
            pending = self.metrics_async(scores)
            if sync:
                res = self.finish(pending)
            else:
                res["pending"] = pending
        res["scores"] = scores
        if "w_mean" in extras:
            for k, ck in (("wi_mean", "classwise_wi"), ("we_mean", "classwise_we")):
                res[k] = self.gather(extra[k])
                res[ck] = {name: res[k][idx] for name, idx in self._class_rows().items()}
        if "wide" in extras:
            local = self._class_rows(local=True)
            for k in ("fused", "image_mu", "event_mu"):
                res[k] = extra[k]
                res["classwise_" + k] = {name: extra[k][idx] for name, idx in local.items()}
        return res

    def _step_overlapped(self, host_inputs: bool) -> Dict[str, object]:
        main = torch.cuda.current_stream(self.device)
        side = self.__dict__.setdefault("_side_stream", None) or torch.cuda.Stream(self.device)
        self._side_stream = side
        packed = self.local_scores(host_inputs)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            packed.record_stream(side)
            scores = self.gather(packed)
            pending = self.metrics_async(scores)
        return {"pending": pending, "scores": scores}

    def _class_rows(self, local: bool = False) -> Dict[str, torch.Tensor]:
        """class key -> device index vector of that class's frames: into the list-order vectors, or (local=True) into this
        rank's compact rows (videos of `self.mine` back to back)."""
        cache = self.__dict__.setdefault("_class_rows_cache", {})
        if local not in cache:
            vids = self.mine if local else range(len(self.lengths))
            if local:
                start = dict(zip(self.mine, np.concatenate([[0], np.cumsum(self.lengths[self.mine])])[:-1])) if len(self.mine) else {}
            else:
                start = dict(enumerate(self.global_off))
            rows: Dict[str, list] = {}
            for v in vids:
                rows.setdefault(self.classes[v], []).append(np.arange(int(start[v]), int(start[v] + self.lengths[v])))
            cache[local] = {k: torch.as_tensor(np.concatenate(a).astype(np.int64), device=self.device) for k, a in rows.items()}
        return cache[local]

    def step_breakdown(self, reps: int = 5, host_inputs: bool = False) -> Dict[str, float]:
        """Where a step's device time goes: forward over my chunks | the single collective (+ re-ordering into list
        order) | ranking (AUC / AP of every subset).  CUDA events on the launching stream, MEDIAN of `reps` steps (a host
        hiccup between two launches shows up as idle stream time in that step's events).  At world > 1 the collective's
        time includes waiting for the slowest rank's forward."""
        marks = []
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            packed = self.local_scores(host_inputs)
            ev[1].record()
            scores = self.gather(packed)
            ev[2].record()
            self.metrics_async(scores)
            ev[3].record()
            marks.append(ev)
        torch.cuda.current_stream(self.device).synchronize()
        names = ("forward_ms", "collective_ms", "ranking_ms")
        med = lambda xs: sorted(xs)[len(xs) // 2]  # noqa: E731
        return {nm: round(med([ev[i].elapsed_time(ev[i + 1]) for ev in marks]), 4) for i, nm in enumerate(names)}

    # bytes moved by a host-input step (for bench.py's e2e block)
    def h2d_bytes(self) -> int:
        if self._ragged is not None:
            return 2 * self._ragged[0].numel() * self._ragged[0].element_size()
        return 2 * self.local_chunks * self.maxlen * self.model.embed_dim * (self._pinned[0].element_size()
                                                                             if self._pinned else 2)

    def d2h_bytes(self) -> int:
        return self.num_subsets * 4 * 8
