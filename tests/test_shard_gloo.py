"""Multi-GPU host logic on CPU: utterance sharding and the all-gather of counts + packed indices, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taste_spokenlm_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _token_counts(n=203, seed=4):
    rng = np.random.default_rng(seed)
    dur = rng.uniform(1, 30, n)
    return np.clip(np.round(2.7 * dur + rng.normal(0, 2, n)), 1, 443).astype(int).tolist()


def test_shards_partition_the_corpus_and_balance_lengths():
    tc = _token_counts()
    for world in (1, 2, 4, 8):
        parts = [shard.shard_indices(tc, world, r) for r in range(world)]
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(len(tc)))                      # disjoint and complete
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
        loads = [sum(tc[i] for i in p) for p in parts]                              # same transcript-length mix
        assert max(loads) - min(loads) <= 0.1 * np.mean(loads) + 64
    owned = shard.shard_indices(tc, 2, 1)
    seen = []
    sizes = []
    for b in shard.batches(owned, tc, 16):
        sizes.append(len(b))
        seen += b.tolist()
    assert seen == owned.tolist()
    assert all(s == 16 for s in sizes[:-1]) and 1 <= sizes[-1] <= 16              # full launches; only the last is short
    lens = [tc[i] for i in owned]
    assert all(lens[i] // 16 <= lens[i + 1] // 16 for i in range(len(lens) - 1))   # bucketed order: neighbours alike
    with pytest.raises(ValueError):
        shard.shard_indices(tc, 2, 2)


def _fake_indices(u, t):
    g = torch.Generator().manual_seed(u)
    return torch.randint(0, 512, (t, 4), generator=g, dtype=torch.int64)


def _worker(rank, world, port, tc, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owned = shard.shard_indices(tc, world, rank)
        ids, res = [], []
        for b in shard.batches(owned, tc, 8):
            for u in b:
                ids.append(int(u))
                res.append(_fake_indices(int(u), tc[u]))
        hdr, flat = shard.pack_results(ids, res)
        got = shard.gather_indices(hdr, flat, 4)
        assert [u for u, _ in got] == list(range(len(tc)))
        for u, t in got:
            assert t.dtype == torch.int16 and t.shape == (tc[u], 4)
            assert torch.equal(t.to(torch.int64), _fake_indices(u, tc[u]))
        torch.save(len(got), os.path.join(out_dir, f"ok{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gather_indices_world2_gloo(tmp_path):
    tc = _token_counts(61, seed=9)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, tc, str(tmp_path)), nprocs=2, join=True)
    assert torch.load(tmp_path / "ok0.pt") == 61 and torch.load(tmp_path / "ok1.pt") == 61


def test_gather_indices_single_process():
    tc = [3, 1, 7]
    hdr, flat = shard.pack_results([2, 0, 1], [_fake_indices(2, 7), _fake_indices(0, 3), _fake_indices(1, 1)])
    got = shard.gather_indices(hdr, flat, 4)
    assert [u for u, _ in got] == [0, 1, 2]
    assert torch.equal(got[2][1].to(torch.int64), _fake_indices(2, 7))
    hdr0, flat0 = shard.pack_results([], [])
    assert shard.gather_indices(hdr0, flat0, 4) == []


def test_gathered_indices_container():
    """The gather returns a lazy, id-sorted view (no per-utterance objects) that still behaves like the list of pairs."""
    hdr, flat = shard.pack_results([5, 2, 9], [_fake_indices(5, 3), _fake_indices(2, 1), _fake_indices(9, 4)])
    got = shard.gather_indices(hdr, flat, 4)
    assert len(got) == 3 and [u for u, _ in got] == [2, 5, 9]
    assert torch.equal(got[1][1].to(torch.int64), _fake_indices(5, 3)) and got[-1][0] == 9
    assert torch.equal(got.lookup(9).to(torch.int64), _fake_indices(9, 4))
    assert [u for u, _ in got[::2]] == [2, 9]
    with pytest.raises(KeyError):
        got.lookup(3)


def test_shard_writer_streams_and_resumes(tmp_path):
    w = shard.ShardWriter(str(tmp_path), rank=1, flush_every=3)
    rng = np.random.default_rng(0)
    ref = {}
    for u in range(7):
        L = int(rng.integers(1, 9))
        li = rng.integers(-1, 512, (L, 4))
        ref[u] = li
        w.add(u, li, rng.integers(0, 32000, L), np.arange(L))
    assert len(w.manifest["files"]) == 2 and sorted(w.done_ids()) == [0, 1, 2, 3, 4, 5]      # 7th still buffered
    # a crash here loses only the unflushed utterance; a restarted rank skips the six on disk
    w2 = shard.ShardWriter(str(tmp_path), rank=1, flush_every=3)
    assert w2.pending(range(9)).tolist() == [6, 7, 8]
    assert w2.is_done(5) and not w2.is_done(6)
    # the batched entry point: padded arrays + lengths
    lens = [4, 2, 5]
    li = rng.integers(-1, 512, (3, 5, 4))
    for r, u in enumerate((6, 7, 8)):
        ref[u] = li[r, : lens[r]]
    w2.add_batch([6, 7, 8], li, rng.integers(0, 32000, (3, 5)), np.tile(np.arange(5), (3, 1)), lens)
    w2.close()
    rows = w2.read_all()
    assert [r["utt_id"] for r in rows] == list(range(9))
    for r in rows:
        assert set(r) == set(shard.ShardWriter.COLUMNS)
        # the reference's row shapes (XV:51-58): [1, L, Q], [1, L], [1], [1, L]
        assert np.array_equal(np.asarray(r["llm_indices"])[0], ref[r["utt_id"]]) and len(r["llm_indices"]) == 1
        assert r["llm_token_lengths"] == [len(r["llm_token_ids"][0])] == [len(r["llm_word_ids"][0])]


def test_shard_writer_output_opens_with_datasets_load_from_disk(tmp_path):
    """Stage-2 training reads the parts with `datasets.load_from_disk` (scripts/run.py:347); the reference writes them
    with `Dataset.from_list(results).save_to_disk` (XV:161-162).  Same rows through both paths must read back equal."""
    datasets = pytest.importorskip("datasets")
    rng = np.random.default_rng(1)
    w = shard.ShardWriter(str(tmp_path / "ours"), rank=0, flush_every=4)
    results = []
    for u in range(10):
        L = int(rng.integers(1, 12))
        li = rng.integers(-1, 512, (L, 4))
        ids, wid = rng.integers(0, 128256, L), np.sort(rng.integers(0, L, L))
        w.add(u, li, ids, wid)
        results.append({                                              # what ExtractVQTrainer.prediction_step appends
            "llm_indices": torch.from_numpy(li)[None].to(torch.int64),
            "llm_token_ids": torch.from_numpy(ids)[None].to(torch.int64),
            "llm_token_lengths": torch.tensor([L], dtype=torch.int32),
            "llm_word_ids": torch.from_numpy(wid)[None].to(torch.int32),
        })
    path = w.finalize()
    ours = datasets.load_from_disk(path)
    ref_dir = str(tmp_path / "ref")
    datasets.Dataset.from_list(results).save_to_disk(ref_dir)
    ref = datasets.load_from_disk(ref_dir)
    assert len(ours) == len(ref) == 10
    for col in ("llm_indices", "llm_token_ids", "llm_token_lengths", "llm_word_ids"):
        assert ours[col] == ref[col] if not hasattr(ours[col], "to_pylist") else list(ours[col]) == list(ref[col]), col
        assert ours.features[col] == ref.features[col], (col, ours.features[col], ref.features[col])
    assert list(ours["utt_id"]) == list(range(10))
