"""Feature ingest: record decoding in the reference's LMDB value format, background staging, and
(GPU) the store-driven path against the in-memory one."""
import io

import numpy as np
import pytest
import torch

from cone_b200 import ingest
from cone_b200.config import EGO4D
from cone_b200.inference import stage_step
from cone_b200.synth import make_dataset


def _stores(ds, eot=False):
    vids = ds.video_ids
    vs = ingest.DictStore({vids[i]: ingest.encode_record(features=ds.videos[i]) for i in range(len(vids))})
    qrec = {}
    for j, q in enumerate(ds.queries):
        hol = {"eot_features": q.cls[None, :]} if (eot and j % 2) else {"cls_features": q.cls}
        qrec[q.query_id] = ingest.encode_record(token_features=q.tokens, **hol)
    return vs, ingest.DictStore(qrec)


def test_records_round_trip_and_reference_reader_semantics():
    ds = make_dataset(EGO4D, 2, [120, 75], [3, 2], seed=4)
    vs, qs = _stores(ds, eot=True)
    for i, vid in enumerate(ds.video_ids):
        got = ingest.decode_video_record(vs.get(vid))
        assert got.dtype == np.float32 and np.array_equal(got, ds.videos[i])
        # the bytes are what the reference's own reader parses (ego4d_mad_dataloader.py:294-299)
        with io.BytesIO(vs.get(vid)) as r:
            assert np.array_equal(np.load(r, allow_pickle=True)["features"], ds.videos[i])
    for q in ds.queries:  # cls_features, or eot_features [1, Dv] squeezed (dataloader:266-271)
        tok, cls = ingest.decode_query_record(qs.get(q.query_id))
        assert np.array_equal(tok, q.tokens) and cls.shape == (EGO4D.v_feat_dim,) and np.array_equal(cls, q.cls)
    with pytest.raises(KeyError):
        ingest.decode_video_record(ingest.encode_record(other=np.zeros(3)))


def test_staged_steps_equal_in_memory_staging_and_surface_errors(tmp_path):
    cfg = EGO4D.replace(eval_bsz=2)
    ds = make_dataset(cfg, 3, [200, 90, 310], [2, 1, 3], seed=6)
    vs, qs = _stores(ds)
    ann = ds.annotations()
    index = {v: i for i, v in enumerate(ds.video_ids)}
    queries = ingest.load_queries(qs, ann, index)
    assert [q.query_id for q in queries] == [q.query_id for q in ds.queries]
    lens = [len(v) for v in ds.videos]
    steps = list(ingest.StagedSteps(cfg, vs, ds.video_ids, lens, queries, max_frames_per_step=300, pin=False))
    assert [s.video_ids for s in steps] == [[0, 1], [2]]
    for s in steps:
        want = stage_step(cfg, ds.videos, ds.queries, s.video_ids, pin=False)
        assert torch.equal(s.frames, want.frames) and torch.equal(s.qb.tokens, want.qb.tokens)
        assert torch.equal(s.qb.cls, want.qb.cls) and torch.equal(s.qb.q_batch, want.qb.q_batch)
        assert s.qb.query_ids == want.qb.query_ids
    # a directory of <key>.npz files is the same store
    for vid in ds.video_ids:
        (tmp_path / (vid + ".npz")).write_bytes(vs.get(vid))
    d = ingest.DirStore(str(tmp_path))
    assert np.array_equal(ingest.decode_video_record(d.get(ds.video_ids[1])), ds.videos[1])
    # a missing record surfaces in the consumer, not in the background thread
    broken = ingest.DictStore({k: v for k, v in vs.records.items() if k != ds.video_ids[2]})
    with pytest.raises(KeyError):
        list(ingest.StagedSteps(cfg, broken, ds.video_ids, lens, queries, max_frames_per_step=300, pin=False))


def test_lmdb_store_fails_loudly_without_lmdb():
    try:
        import lmdb  # noqa: F401
        pytest.skip("lmdb is installed here")
    except ImportError:
        with pytest.raises(ImportError):
            ingest.LmdbStore("/nonexistent")


@pytest.mark.gpu
def test_ground_store_equals_in_memory_path():
    from cone_b200.engine import ConeEngine
    from cone_b200.inference import ground_dataset
    from cone_b200.weights import init_state_dict
    cfg = EGO4D.replace(eval_bsz=2)
    ds = make_dataset(cfg, 3, [400, 90, 310], [2, 2, 2], seed=8)
    vs, qs = _stores(ds, eot=True)
    eng = ConeEngine(cfg, init_state_dict(cfg, 2), device="cuda:0", precision="fp32", workspace_bytes=1 << 30)
    a = ingest.ground_store(eng, vs, qs, ds.annotations(), max_frames_per_step=500)
    b = ground_dataset(eng, ds.videos, ds.queries, max_frames_per_step=500, full=False)
    assert a.keys() == b.keys()
    for k in a:
        assert a[k] == b[k]


def test_staged_steps_producer_exits_when_the_consumer_stops_early():
    """a consumer that breaks out of the loop must not leave the producer blocked on queue.put."""
    cfg = EGO4D.replace(eval_bsz=2)
    ds = make_dataset(cfg, 6, [100] * 6, 1, seed=9)
    vs, qs = _stores(ds)
    queries = ingest.load_queries(qs, ds.annotations(), {v: i for i, v in enumerate(ds.video_ids)})
    steps = ingest.StagedSteps(cfg, vs, ds.video_ids, [100] * 6, queries, max_frames_per_step=100, depth=1, pin=False)
    for s in steps:
        break  # abandon after the first of six steps
    steps._thread.join(timeout=5.0)
    assert not steps._thread.is_alive()
    with ingest.StagedSteps(cfg, vs, ds.video_ids, [100] * 6, queries, max_frames_per_step=100, depth=1, pin=False) as st2:
        pass  # never iterated
    assert not st2._thread.is_alive()


def test_records_are_loaded_without_pickle():
    import pickle

    class Evil:
        def __reduce__(self):
            return (pytest.fail, ("pickle payload executed",))
    rec = ingest.encode_record(features=np.array([Evil()], dtype=object))
    with pytest.raises((ValueError, pickle.UnpicklingError)):
        ingest.decode_video_record(rec)


def test_eval_batch_ids_follow_true_dataset_indices():
    """q_batch = dataset index // eval_bsz even when a step's queries are not contiguous in the dataset."""
    from cone_b200.engine import pack_queries
    cfg = EGO4D.replace(eval_bsz=4)
    ds = make_dataset(cfg, 3, [100, 120, 90], [3, 3, 3], seed=2)
    # interleave the annotations: dataset order v0 v1 v2 v0 v1 v2 ...
    inter = [ds.queries[3 * v + j] for j in range(3) for v in range(3)]
    step = stage_step(cfg, ds.videos, inter, [0, 2], pin=False)  # queries of videos 0 and 2: dataset idx 0,2,3,5,6,8
    by_id = dict(zip(step.qb.query_ids, step.qb.q_batch.tolist()))
    want = {}
    for i, q in enumerate(inter):
        if q.video_idx in (0, 2):
            want[q.query_id] = i // 4
    dense = {b: n for n, b in enumerate(sorted(set(want.values())))}
    assert by_id == {k: dense[b] for k, b in want.items()}
    assert step.qb.n_batches == len(dense) == 3
    # contiguous default unchanged
    qb = pack_queries(cfg, [100, 120, 90], ds.queries, first_dataset_index=6)
    assert sorted(set(qb.q_batch.tolist())) == [0, 1, 2] and qb.n_batches == 3
