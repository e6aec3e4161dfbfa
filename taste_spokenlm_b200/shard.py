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
    """Cut a rank's index list into batches that never straddle a length bucket boundary by more than one bucket."""
    tc = np.asarray(token_counts, dtype=np.int64)
    start = 0
    n = len(owned)
    while start < n:
        end = min(start + batch_size, n)
        b0 = tc[owned[start]] // bucket_width
        # keep the batch within two adjacent buckets so padding stays < 2 * bucket_width tokens per row
        while end > start + 1 and tc[owned[end - 1]] // bucket_width > b0 + 1:
            end -= 1
        yield owned[start:end]
        start = end


def pack_results(utt_ids: Sequence[int], indices: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-utterance `[T_b, Q]` index tensors -> (header int32 [n, 2] = (utterance id, T_b), packed int16 [sum T_b, Q])."""
    hdr = torch.tensor([[int(u), int(t.shape[0])] for u, t in zip(utt_ids, indices)], dtype=torch.int32).reshape(-1, 2)
    if len(indices):
        flat = torch.cat([t.reshape(t.shape[0], -1) for t in indices]).to(torch.int16)
    else:
        flat = torch.zeros(0, 0, dtype=torch.int16)
    return hdr, flat


def gather_indices(hdr: torch.Tensor, flat: torch.Tensor, num_q: int, device=None) -> List[Tuple[int, torch.Tensor]]:
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


def _unpack(parts) -> List[Tuple[int, torch.Tensor]]:
    out = []
    for hdr, flat in parts:
        off = 0
        for u, t in hdr.tolist():
            out.append((int(u), flat[off: off + t]))
            off += t
    out.sort(key=lambda x: x[0])
    return out


def tokenize_corpus(engine, corpus, world_size: int, rank: int, batch_size: int = 64, gather: bool = True):
    """Tokenize this rank's shard of `corpus` and (optionally) gather everyone's indices.

    `corpus`: object with `token_counts` (list[int]) and `load(indices) -> dict` returning device tensors
    `wav [B, N] f32`, `n_samples [B] i32`, `ids [B, Tmax] i64`, `wid [B, Tmax] i32`, `lengths_host np[B]`, where
    `Tmax == lengths_host.max()`.  `engine`: a packed `TowerEngine`.
    """
    owned = shard_indices(corpus.token_counts, world_size, rank)
    utt_ids, results = [], []
    for b in batches(owned, corpus.token_counts, batch_size):
        batch = corpus.load(b)
        _, idx = engine.tokenize_device(batch["wav"], batch["n_samples"], batch["ids"], batch["wid"],
                                        batch["lengths_host"], want_quantized=False)
        for row, u in enumerate(b):
            T = int(batch["lengths_host"][row])
            utt_ids.append(int(u))
            results.append(idx[row, :T])
    hdr, flat = pack_results(utt_ids, results)
    num_q = engine.cfg.num_quantizers
    if not gather:
        return _unpack([(hdr, flat.cpu().reshape(-1, num_q))])
    return gather_indices(hdr.to(engine.device), flat.to(engine.device), num_q, engine.device)


class ShardWriter:
    """Streaming, resumable result writer for the corpus job.

    The reference buffers every result in RAM and writes one HF dataset per rank at the very end
    (`Dataset.from_list(self.results).save_to_disk(f"{output_dir}/part-{LOCAL_RANK}")`, XV:70, XV:161-162): a crash loses
    the rank's whole shard.  This writer flushes an Arrow IPC file every `flush_every` utterances under
    `<out_dir>/part-<rank>/` with the reference's column names (`llm_indices`, `llm_token_ids`, `llm_token_lengths`,
    `llm_word_ids`; XV:51-70) plus `utt_id`, and keeps `manifest.json` (files + utterance ids done) so a restarted rank
    skips what is already on disk.
    """

    COLUMNS = ("utt_id", "llm_indices", "llm_token_ids", "llm_token_lengths", "llm_word_ids")

    def __init__(self, out_dir: str, rank: int, flush_every: int = 1024):
        import json
        import os
        self.dir = os.path.join(out_dir, f"part-{rank}")
        os.makedirs(self.dir, exist_ok=True)
        self.flush_every = flush_every
        self._manifest_path = os.path.join(self.dir, "manifest.json")
        self.manifest = {"files": [], "done": []}
        if os.path.exists(self._manifest_path):
            with open(self._manifest_path) as f:
                self.manifest = json.load(f)
        self._done = set(self.manifest["done"])
        self._rows = []

    def is_done(self, utt_id: int) -> bool:
        return int(utt_id) in self._done

    def pending(self, utt_ids) -> np.ndarray:
        """The subset of `utt_ids` that still has to be tokenized (order preserved)."""
        return np.asarray([u for u in utt_ids if int(u) not in self._done], dtype=np.int64)

    def add(self, utt_id: int, llm_indices, llm_token_ids, llm_word_ids) -> None:
        """One utterance: llm_indices [L, Q] (ints, -1 on non-word-start tokens), llm_token_ids [L], llm_word_ids [L]."""
        li = np.asarray(llm_indices, dtype=np.int64)
        self._rows.append((int(utt_id), li.tolist(), np.asarray(llm_token_ids, dtype=np.int64).tolist(), int(li.shape[0]),
                           np.asarray(llm_word_ids, dtype=np.int64).tolist()))
        if len(self._rows) >= self.flush_every:
            self.flush()

    def flush(self) -> None:
        import json
        import os
        import pyarrow as pa
        if not self._rows:
            return
        cols = list(zip(*self._rows))
        table = pa.table({name: list(col) for name, col in zip(self.COLUMNS, cols)})
        name = f"data-{len(self.manifest['files']):05d}.arrow"
        tmp = os.path.join(self.dir, name + ".tmp")
        with pa.OSFile(tmp, "wb") as sink, pa.ipc.new_file(sink, table.schema) as w:
            w.write_table(table)
        os.replace(tmp, os.path.join(self.dir, name))              # the file exists completely or not at all
        self.manifest["files"].append(name)
        self.manifest["done"].extend(int(u) for u in cols[0])
        self._done.update(int(u) for u in cols[0])
        with open(self._manifest_path + ".tmp", "w") as f:
            json.dump(self.manifest, f)
        os.replace(self._manifest_path + ".tmp", self._manifest_path)
        self._rows = []

    def close(self) -> None:
        self.flush()

    def read_all(self):
        """All rows written so far as a list of dicts (for tests / small shards)."""
        import os
        import pyarrow as pa
        out = []
        for name in self.manifest["files"]:
            with pa.memory_map(os.path.join(self.dir, name)) as src:
                out.extend(pa.ipc.open_file(src).read_all().to_pylist())
        return out
