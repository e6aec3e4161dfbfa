"""Utterance sharding and the result gather for corpus-scale tokenization (BASELINE config 4; reference workload:
scripts/extract_vq_for_stage2_training.py, XV:137-165 — one process per GPU, DistributedSampler-style sharding, each
rank keeps its own results, no collective on the data path).

Utterances are independent, and every utterance costs one full 30 s encoder window regardless of its
audio length, so the only cost that varies is the transcript length.  `shard_indices` therefore sorts the
corpus into transcript-length buckets (which also keeps the padded `[B, Tmax]` token batches tight: the tower requires
the padded width to equal the longest transcript of the batch, MT:181-184) and deals every bucket round-robin to the
ranks, so each rank sees the same length mix.  `gather_indices` is the path's only collective: one all-gather of the
per-rank token counts followed by one all-gather of the packed int16 indices (NCCL over NVLink on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import Iterator, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(token_counts: Sequence[int], world_size: int, rank: int, bucket_width: int = 16) -> np.ndarray:
    """Indices (into the corpus) owned by `rank`, ordered bucket by bucket (short transcripts first)."""
    tc = np.asarray(token_counts, dtype=np.int64)
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    order = np.lexsort((np.arange(len(tc)), tc // bucket_width))      # stable: bucket, then corpus order
    return order[rank::world_size]


def batches(owned: np.ndarray, token_counts: Sequence[int], batch_size: int, bucket_width: int = 16
            ) -> Iterator[np.ndarray]:
    """Cut a rank's (length-bucketed) index list into FULL batches, in order; only the last one may be short.

    Every utterance costs one full encoder window whatever its transcript length and the aggregator runs on packed
    rows, so a batch's cost does not depend on how its transcript lengths mix: what matters is that every launch has
    `batch_size` windows and that all ranks have the same number of batches (round 2: cutting at bucket boundaries left
    ~7 % of the batches short and the ranks up to a second apart at the final gather of a 100 k-utterance job).  The
    bucketed order still keeps neighbouring lengths together, i.e. the padded `[B, Tmax]` id tensors tight."""
    for start in range(0, len(owned), batch_size):
        yield owned[start: start + batch_size]


def pack_results(utt_ids: Sequence[int], indices: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-utterance `[T_b, Q]` index tensors -> (header int32 [n, 2] = (utterance id, T_b), packed int16 [sum T_b, Q])."""
    hdr = torch.tensor([[int(u), int(t.shape[0])] for u, t in zip(utt_ids, indices)], dtype=torch.int32).reshape(-1, 2)
    if len(indices):
        flat = torch.cat([t.reshape(t.shape[0], -1) for t in indices]).to(torch.int16)
    else:
        flat = torch.zeros(0, 0, dtype=torch.int16)
    return hdr, flat


def gather_indices(hdr: torch.Tensor, flat: torch.Tensor, num_q: int, device=None) -> GatheredIndices:
    """All ranks receive every (utterance id, [T, Q] int16 indices), sorted by utterance id.

    Two collectives: all_gather of the (n_utts, n_rows) counts, then all_gather of the padded header / index buffers.
    Codebook size 512 fits int16, which quarters the bytes moved versus the int64 the model returns.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return _unpack([(hdr.cpu(), flat.cpu().reshape(-1, num_q))])
    world = dist.get_world_size()
    dev = device if device is not None else hdr.device
    counts = torch.tensor([hdr.shape[0], flat.shape[0]], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    max_u = int(max(int(c[0]) for c in all_counts))
    max_r = int(max(int(c[1]) for c in all_counts))
    hbuf = torch.zeros(max_u, 2, dtype=torch.int32, device=dev)
    hbuf[: hdr.shape[0]] = hdr.to(dev)
    fbuf = torch.zeros(max_r, num_q, dtype=torch.int16, device=dev)
    if flat.numel():
        fbuf[: flat.shape[0]] = flat.to(dev).reshape(-1, num_q)
    hall = [torch.zeros_like(hbuf) for _ in range(world)]
    fall = [torch.zeros_like(fbuf) for _ in range(world)]
    dist.all_gather(hall, hbuf)
    # moved as raw bytes: gloo has no int16 collectives, and the byte count is what matters on NVLink anyway
    dist.all_gather([t.view(torch.uint8) for t in fall], fbuf.view(torch.uint8))
    parts = []
    for r in range(world):
        nu, nr = int(all_counts[r][0]), int(all_counts[r][1])
        parts.append((hall[r][:nu].cpu(), fall[r][:nr].cpu()))
    return _unpack(parts)


class GatheredIndices:
    """The job's results, sorted by utterance id, without one Python object per utterance: `ids[i]`, and the `[T_i, Q]`
    int16 indices of that utterance as a view of one flat tensor.  Behaves like the list of `(utterance id, indices)`
    pairs it replaces (len / iteration / integer indexing / == []).  With 100 k utterances per job, building that list on
    every rank was 1.5 s of a 30 s job."""

    def __init__(self, ids: np.ndarray, lens: np.ndarray, flat: torch.Tensor):
        order = np.argsort(ids, kind="stable")
        starts = np.zeros(len(ids) + 1, dtype=np.int64)
        np.cumsum(lens, out=starts[1:])
        self.ids = ids[order]
        self._start = starts[:-1][order]
        self._len = lens[order]
        self.flat = flat

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        s = int(self._start[i])
        return int(self.ids[i]), self.flat[s: s + int(self._len[i])]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __eq__(self, other):
        return list(self) == other if isinstance(other, list) else NotImplemented

    def lookup(self, utt_id: int) -> torch.Tensor:
        i = int(np.searchsorted(self.ids, utt_id))
        if i >= len(self.ids) or int(self.ids[i]) != int(utt_id):
            raise KeyError(utt_id)
        return self[i][1]


def _unpack(parts) -> GatheredIndices:
    """[(header [n, 2], flat [rows, Q])] per rank -> the results sorted by utterance id (views of one tensor)."""
    parts = [(h, f) for h, f in parts if h.shape[0] > 0]
    if not parts:
        return GatheredIndices(np.zeros(0, np.int64), np.zeros(0, np.int64), torch.zeros(0, 0, dtype=torch.int16))
    hdr = torch.cat([h for h, _ in parts]).numpy().astype(np.int64)
    flat = torch.cat([f for _, f in parts])
    return GatheredIndices(hdr[:, 0], hdr[:, 1], flat)


class _Slot:
    """One in-flight batch of the corpus pipeline: pinned host staging + device buffers + the events that order them."""

    def __init__(self, device, batch_size: int, wav_stride: int, max_t: int, max_l: int, num_q: int):
        self.wav_h = torch.empty(batch_size, wav_stride, dtype=torch.float32).pin_memory()
        self.ns_h = torch.zeros(batch_size, dtype=torch.int32).pin_memory()
        self.ids_h = torch.zeros(batch_size, max_t, dtype=torch.int64).pin_memory()
        self.wid_h = torch.zeros(batch_size, max_t, dtype=torch.int32).pin_memory()
        self.lwid_h = torch.zeros(batch_size, max_l, dtype=torch.int32).pin_memory()
        self.len_h = torch.zeros(2, batch_size, dtype=torch.int32).pin_memory()          # asr lengths, llm lengths
        self.wav_d = torch.empty(batch_size, wav_stride, dtype=torch.float32, device=device)
        self.ns_d = torch.zeros(batch_size, dtype=torch.int32, device=device)
        self.ids_d = torch.zeros(batch_size, max_t, dtype=torch.int64, device=device)
        self.wid_d = torch.zeros(batch_size, max_t, dtype=torch.int32, device=device)
        self.lwid_d = torch.zeros(batch_size, max_l, dtype=torch.int32, device=device)
        self.len_d = torch.zeros(2, batch_size, dtype=torch.int32, device=device)
        self.out_h = torch.empty(batch_size * (max_t + max_l) * num_q, dtype=torch.int16).pin_memory()
        self.ready = torch.cuda.Event()          # H2D of this slot's inputs finished (copy stream)
        self.consumed = torch.cuda.Event()       # the kernels that read the device buffers finished (compute stream)
        self.result = torch.cuda.Event()         # D2H of this slot's indices finished (compute stream)
        self.meta = None


def tokenize_corpus(engine, corpus, world_size: int, rank: int, batch_size: int = 64, gather: bool = True,
                    writer: "ShardWriter" = None, timings: dict = None, map_llm: bool = None):
    """Tokenize this rank's shard of `corpus` and (optionally) gather everyone's indices.

    Two corpus protocols:

    * host-resident (the corpus job, XV:137-162) — `corpus.fetch_host(indices, slot) -> meta` fills the pinned staging
      tensors of `slot` (`wav_h [B, N] f32`, `ns_h`, `ids_h`, `wid_h`, and for the llm mapping `lwid_h`, `len_h[1]`) and
      returns `{"B", "lengths_host", "max_n", "T", ["L", "llm_lengths_host", "llm_ids"]}`.  Batches run through a
      three-slot pipeline: a worker thread fills slot i+1 and enqueues its H2D on a copy stream while the compute
      stream runs `tokenize_device` (+ `map_to_llm_tokens`) on slot i and the host hands batch i-1 to the writer; every
      batch makes ONE device-to-host copy (int16 indices into pinned memory), consumed one batch later, so the host
      never waits on the GPU except for back-pressure.  Results go to `writer.add_batch` (vectorised) and/or are kept for the final gather.
    * device-resident (tests, small jobs) — `corpus.load(indices) -> dict` of device tensors, run synchronously.

    `timings` (optional dict) receives the per-stage host times in ms: fetch (worker thread), stall (compute thread
    waiting for the worker), result_wait (waiting for a batch's D2H), writer, gather, and the batch count.
    """
    owned = shard_indices(corpus.token_counts, world_size, rank)
    if writer is not None:
        keep = ~np.isin(owned, np.fromiter(writer.done_ids(), dtype=np.int64, count=-1))
        owned = owned[keep]
    num_q = engine.cfg.num_quantizers
    tm = timings if timings is not None else {}
    for k in ("fetch_ms", "stall_ms", "result_wait_ms", "writer_ms", "gather_ms"):
        tm.setdefault(k, 0.0)
    tm["batches"] = 0
    if not hasattr(corpus, "fetch_host"):
        utt_ids, results = [], []
        for b in batches(owned, corpus.token_counts, batch_size):
            batch = corpus.load(b)
            _, idx = engine.tokenize_device(batch["wav"], batch["n_samples"], batch["ids"], batch["wid"],
                                            batch["lengths_host"], want_quantized=False)
            idx16 = idx.to(torch.int16).cpu()                      # one D2H per batch
            for row, u in enumerate(b):
                utt_ids.append(int(u))
                results.append(idx16[row, : int(batch["lengths_host"][row])])
            tm["batches"] += 1
        hdr, flat = pack_results(utt_ids, results)
    else:
        if writer is not None and map_llm is False:
            raise ValueError("the writer stores llm-token-aligned rows (XV:51-58): map_llm cannot be False with a writer")
        with torch.cuda.device(engine.device):          # streams, events and pinned buffers of the pipeline live there
            hdr, flat = _run_pipeline(engine, corpus, owned, batch_size, writer, tm, num_q,
                                      bool(map_llm) if map_llm is not None else writer is not None, keep_results=gather)
    if writer is not None:
        t0 = _now()
        writer.flush()
        tm["writer_ms"] += (_now() - t0) * 1e3
    if not gather:
        return _unpack([(hdr, flat.cpu().reshape(-1, num_q))])
    t0 = _now()
    out = gather_indices(hdr.to(engine.device), flat.to(engine.device), num_q, engine.device)
    tm["gather_ms"] += (_now() - t0) * 1e3
    return out


def _now() -> float:
    import time
    return time.perf_counter()


N_SLOTS = 3      # batch i computes, i + 1 is being staged, i - 1 waits for its host-side hand-over


def _run_pipeline(engine, corpus, owned, batch_size, writer, tm, num_q, map_llm, keep_results):
    from concurrent.futures import ThreadPoolExecutor
    dev = engine.device
    blist = list(batches(owned, corpus.token_counts, batch_size))
    if not blist:
        return torch.zeros(0, 2, dtype=torch.int32), torch.zeros(0, num_q, dtype=torch.int16)
    max_t = int(max(corpus.token_counts[i] for i in owned))
    max_l = int(getattr(corpus, "max_llm_tokens", max_t)) if map_llm else 1
    slots = [_Slot(dev, batch_size, int(getattr(corpus, "wav_stride", 480000)), max_t, max_l, num_q) for _ in range(N_SLOTS)]
    copy_stream = torch.cuda.Stream(device=dev)
    compute = torch.cuda.current_stream(dev)
    for s in slots:
        s.consumed.record(compute)
        s.result.record(compute)

    def prepare(i):
        torch.cuda.set_device(dev)
        s = slots[i % N_SLOTS]
        s.result.synchronize()               # the host has to be done with this slot's previous result as well ...
        s.consumed.synchronize()             # ... and the kernels with its device buffers (back-pressure)
        t0 = _now()
        meta = corpus.fetch_host(blist[i], s)
        tm["fetch_ms"] += (_now() - t0) * 1e3
        B, T, n = meta["B"], meta["T"], int(meta["max_n"])
        with torch.cuda.stream(copy_stream):
            s.wav_d[:B, :n].copy_(s.wav_h[:B, :n], non_blocking=True)         # ragged: only the audio that exists
            s.ns_d[:B].copy_(s.ns_h[:B], non_blocking=True)
            s.ids_d[:B, :T].copy_(s.ids_h[:B, :T], non_blocking=True)
            s.wid_d[:B, :T].copy_(s.wid_h[:B, :T], non_blocking=True)
            if map_llm:
                L = meta["L"]
                s.lwid_d[:B, :L].copy_(s.lwid_h[:B, :L], non_blocking=True)
                s.len_d[:, :B].copy_(s.len_h[:, :B], non_blocking=True)
            s.ready.record(copy_stream)
        meta["h2d_bytes"] = 4 * B * n + 12 * B * T + 4 * B + (4 * B * meta["L"] + 8 * B if map_llm else 0)
        return meta

    res_ids, res_lens, res_rows = [], [], []       # per batch: utterance ids, token counts, ragged [sum T, Q] int16 rows
    tm.setdefault("h2d_bytes", 0)
    tm.setdefault("d2h_bytes", 0)

    def finish(i, meta):
        """Host side of batch i, one batch behind the GPU: wait for its D2H, hand the rows to the writer / the gather."""
        s = slots[i % N_SLOTS]
        t0 = _now()
        s.result.synchronize()
        tm["result_wait_ms"] += (_now() - t0) * 1e3
        B = meta["B"]
        W = meta["L"] if map_llm else meta["T"]
        out = s.out_h[: B * W * num_q].view(B, W, num_q).numpy()
        if keep_results or writer is None:
            asr = out if not map_llm else None
            if map_llm:            # the gather carries asr-token indices: first token of every word <-> llm word starts
                asr = s.out_h[B * W * num_q: B * W * num_q + B * meta["T"] * num_q].view(B, meta["T"], num_q).numpy()
            lens = np.asarray(meta["lengths_host"], dtype=np.int64)
            res_ids.append(np.asarray(blist[i], dtype=np.int64))
            res_lens.append(lens)
            res_rows.append(asr[np.arange(asr.shape[1])[None, :] < lens[:, None]])    # copies out of the staging buffer
        if writer is not None:
            t0 = _now()
            writer.add_batch(blist[i], out, meta["llm_ids"], s.lwid_h[:B, :W].numpy(), meta["llm_lengths_host"])
            tm["writer_ms"] += (_now() - t0) * 1e3

    from collections import deque
    with ThreadPoolExecutor(max_workers=1) as pool:
        fut = pool.submit(prepare, 0)
        inflight = deque()                   # launched, host side not yet handed over
        for i in range(len(blist)):
            t0 = _now()
            meta = fut.result()
            tm["stall_ms"] += (_now() - t0) * 1e3
            # slot (i + 1) % 3 was last used by batch i - 2, whose kernels finished long ago: handing it over does not
            # block, and batch i - 1 keeps the GPU busy while batch i's launches are enqueued below
            while inflight and inflight[0][0] <= i - (N_SLOTS - 1):
                finish(*inflight.popleft())
            if i + 1 < len(blist):
                fut = pool.submit(prepare, i + 1)
            s = slots[i % N_SLOTS]
            B, T = meta["B"], meta["T"]
            compute.wait_event(s.ready)
            _, idx = engine.tokenize_device(s.wav_d[:B], s.ns_d[:B], s.ids_d[:B, :T].contiguous(),
                                            s.wid_d[:B, :T].contiguous(), meta["lengths_host"], want_quantized=False)
            if map_llm:
                L = meta["L"]
                llm = engine.map_to_llm_tokens(idx, s.wid_d[:B, :T], s.len_d[0, :B], s.lwid_d[:B, :L].contiguous(),
                                               s.len_d[1, :B])
                n_out = B * L * num_q
                s.out_h[:n_out].copy_(llm.to(torch.int16).reshape(-1), non_blocking=True)
                if keep_results or writer is None:
                    s.out_h[n_out: n_out + B * T * num_q].copy_(idx.to(torch.int16).reshape(-1), non_blocking=True)
                    n_out += B * T * num_q
            else:
                n_out = B * T * num_q
                s.out_h[:n_out].copy_(idx.to(torch.int16).reshape(-1), non_blocking=True)
            s.consumed.record(compute)
            s.result.record(compute)
            tm["h2d_bytes"] += meta["h2d_bytes"]
            tm["d2h_bytes"] += 2 * n_out
            tm["batches"] += 1
            inflight.append((i, meta))
        while inflight:
            finish(*inflight.popleft())
    if not res_ids:
        return torch.zeros(0, 2, dtype=torch.int32), torch.zeros(0, num_q, dtype=torch.int16)
    hdr = torch.from_numpy(np.stack([np.concatenate(res_ids), np.concatenate(res_lens)], axis=1).astype(np.int32))
    return hdr, torch.from_numpy(np.concatenate(res_rows).reshape(-1, num_q))


class ShardWriter:
    """Streaming, resumable result writer for the corpus job.

    The reference buffers every result in RAM and writes one HF dataset per rank at the very end
    (`Dataset.from_list(self.results).save_to_disk(f"{output_dir}/part-{LOCAL_RANK}")`, XV:70, XV:161-162): a crash loses
    the rank's whole shard.  This writer flushes one Arrow file every `flush_every` utterances under
    `<out_dir>/part-<rank>/`, in the layout `datasets.load_from_disk` reads (Arrow IPC *stream* files named
    `data-XXXXX-of-NNNNN.arrow` + `state.json` + `dataset_info.json`, written by `finalize()`), with the reference's
    columns and row shapes (XV:51-58): `llm_indices [1, L, Q] int64`, `llm_token_ids [1, L] int64`,
    `llm_token_lengths [1] int32`, `llm_word_ids [1, L] int32`, plus `utt_id`.  `manifest.json` lists the finished files;
    a restarted rank reads the `utt_id` column of those files and skips what is already on disk.
    Rows are appended a batch at a time as numpy arrays and turned into Arrow list arrays from flat buffers (no per-row
    Python objects).
    """

    COLUMNS = ("utt_id", "llm_indices", "llm_token_ids", "llm_token_lengths", "llm_word_ids")

    def __init__(self, out_dir: str, rank: int, flush_every: int = 1024):
        import json
        import os
        self.dir = os.path.join(out_dir, f"part-{rank}")
        os.makedirs(self.dir, exist_ok=True)
        self.flush_every = flush_every
        self._manifest_path = os.path.join(self.dir, "manifest.json")
        self.manifest = {"files": [], "rows": []}
        if os.path.exists(self._manifest_path):
            with open(self._manifest_path) as f:
                self.manifest = json.load(f)
            self.manifest.setdefault("rows", [None] * len(self.manifest["files"]))
        self._done = set()
        for name in self.manifest["files"]:
            self._done.update(int(u) for u in self._read_table(name).column("utt_id").to_numpy())
        self._pending = []            # (utt_ids [n], idx [n, L, Q], ids [n, L], wid [n, L], lens [n]) per batch
        self._n_pending = 0

    # ---- bookkeeping ----
    def is_done(self, utt_id: int) -> bool:
        return int(utt_id) in self._done

    def done_ids(self):
        return self._done

    def pending(self, utt_ids) -> np.ndarray:
        """The subset of `utt_ids` that still has to be tokenized (order preserved)."""
        return np.asarray([u for u in utt_ids if int(u) not in self._done], dtype=np.int64)

    # ---- adding rows ----
    def add(self, utt_id: int, llm_indices, llm_token_ids, llm_word_ids) -> None:
        """One utterance: llm_indices [L, Q] (ints, -1 on non-word-start tokens), llm_token_ids [L], llm_word_ids [L]."""
        li = np.asarray(llm_indices)
        L = li.shape[0]
        self.add_batch([utt_id], li[None], np.asarray(llm_token_ids)[None], np.asarray(llm_word_ids)[None], [L])

    def add_batch(self, utt_ids, llm_indices, llm_token_ids, llm_word_ids, llm_lengths) -> None:
        """A batch: llm_indices [n, Lpad, Q], llm_token_ids / llm_word_ids [n, Lpad] (padded), llm_lengths [n].
        The arrays are copied (callers reuse their staging buffers)."""
        lens = np.asarray(llm_lengths, dtype=np.int64)
        n = len(lens)
        self._pending.append((np.asarray(utt_ids, dtype=np.int64).copy(),
                              np.array(llm_indices[:n], dtype=np.int64), np.array(llm_token_ids[:n], dtype=np.int64),
                              np.array(llm_word_ids[:n], dtype=np.int32), lens.copy()))
        self._n_pending += n
        if self._n_pending >= self.flush_every:
            self.flush()

    @staticmethod
    def _ragged(mat: np.ndarray, lens: np.ndarray) -> np.ndarray:
        """Rows of a padded [n, Lpad, ...] array cut to their lengths and concatenated."""
        mask = np.arange(mat.shape[1])[None, :] < lens[:, None]
        return mat[mask]

    def _table(self):
        import pyarrow as pa
        utt = np.concatenate([p[0] for p in self._pending])
        lens = np.concatenate([p[4] for p in self._pending])
        idx = np.concatenate([self._ragged(p[1], p[4]) for p in self._pending])          # [sum L, Q]
        ids = np.concatenate([self._ragged(p[2], p[4]) for p in self._pending])          # [sum L]
        wid = np.concatenate([self._ragged(p[3], p[4]) for p in self._pending])
        n, Q = len(lens), idx.shape[1] if idx.ndim == 2 else 1
        row_off = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(lens, out=row_off[1:])
        one = np.arange(n + 1, dtype=np.int32)                                           # the leading [1, ...] dimension

        def lists(values, offsets):
            return pa.ListArray.from_arrays(pa.array(offsets, type=pa.int32()), values)

        inner = lists(pa.array(idx.reshape(-1), type=pa.int64()), np.arange(idx.shape[0] + 1, dtype=np.int32) * Q)
        cols = {
            "utt_id": pa.array(utt, type=pa.int64()),
            "llm_indices": lists(lists(inner, row_off), one),                            # [1, L, Q]
            "llm_token_ids": lists(lists(pa.array(ids, type=pa.int64()), row_off), one),   # [1, L]
            "llm_token_lengths": lists(pa.array(lens.astype(np.int32), type=pa.int32()), one),   # [1]
            "llm_word_ids": lists(lists(pa.array(wid, type=pa.int32()), row_off), one),    # [1, L]
        }
        return pa.table(cols), utt

    def flush(self) -> None:
        import json
        import os
        import pyarrow as pa
        if not self._pending:
            return
        table, utt = self._table()
        name = f"part-{len(self.manifest['files']):05d}.arrow"
        tmp = os.path.join(self.dir, name + ".tmp")
        with pa.OSFile(tmp, "wb") as sink, pa.ipc.new_stream(sink, table.schema) as w:
            w.write_table(table)
        os.replace(tmp, os.path.join(self.dir, name))              # the file exists completely or not at all
        self.manifest["files"].append(name)
        self.manifest["rows"].append(int(len(utt)))
        self._done.update(int(u) for u in utt)
        with open(self._manifest_path + ".tmp", "w") as f:
            json.dump(self.manifest, f)
        os.replace(self._manifest_path + ".tmp", self._manifest_path)
        self._pending, self._n_pending = [], 0

    def close(self) -> None:
        self.flush()

    def finalize(self) -> str:
        """Make `<out_dir>/part-<rank>` a directory `datasets.load_from_disk` opens (what stage-2 training reads,
        scripts/run.py:347): `state.json` listing the data files and `dataset_info.json` with the features.  Uses the
        `datasets` package (a dependency of the reference's own script, XV:4) only for the feature description."""
        import json
        import os
        self.flush()
        files = list(self.manifest["files"])
        if not files:
            raise RuntimeError("nothing written")
        from datasets import Features
        schema = self._read_table(files[0]).schema
        info = {"citation": "", "description": "", "features": Features.from_arrow_schema(schema).to_dict(),
                "homepage": "", "license": ""}
        state = {"_data_files": [{"filename": f} for f in files], "_fingerprint": f"taste-b200-{len(files):05d}",
                 "_format_columns": None, "_format_kwargs": {}, "_format_type": None, "_output_all_columns": False,
                 "_split": None}
        with open(os.path.join(self.dir, "dataset_info.json"), "w") as f:
            json.dump(info, f, indent=2)
        with open(os.path.join(self.dir, "state.json"), "w") as f:
            json.dump(state, f, indent=2)
        return self.dir

    def _read_table(self, name):
        import os
        import pyarrow as pa
        with pa.memory_map(os.path.join(self.dir, name)) as src:
            return pa.ipc.open_stream(src).read_all()

    def read_all(self):
        """All rows written so far as a list of dicts (for tests / small shards)."""
        out = []
        for name in self.manifest["files"]:
            out.extend(self._read_table(name).to_pylist())
        return out
