"""The columnar store (engine/store.py, format 2) without a GPU: column files, the commit protocol (column tails first,
header replaced last, tails truncated to the committed count on open), failure roll-back, shard row ranges, lazy rows.
The device table is replaced by a recording fake; search itself is covered by tests/test_services_gpu.py."""
import importlib
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeIndex:
    """Stands in for engine.index.VectorIndex: keeps what was appended."""
    fail_next = False

    def __init__(self, dim, device=0, capacity=0, keep_f32=False):
        self.dim, self.keep_f32 = dim, keep_f32
        self.vecs = np.zeros((0, dim), np.float32)
        self.levels = np.zeros((0,), np.uint8)
        self.closed = False

    def append(self, vecs, levels=None):
        if FakeIndex.fail_next:
            FakeIndex.fail_next = False
            raise RuntimeError("injected device failure")
        v = np.asarray(vecs)
        if v.dtype == np.uint16:
            v = (v.astype(np.uint32) << 16).view(np.float32)
        self.vecs = np.concatenate([self.vecs, v.astype(np.float32)])
        self.levels = np.concatenate([self.levels, np.asarray(levels, np.uint8)])

    def __len__(self):
        return len(self.vecs)

    def close(self):
        self.closed = True


@pytest.fixture()
def store(monkeypatch):
    S = importlib.import_module("rag-project-icd10_b200.engine.store")
    N = importlib.import_module("rag-project-icd10_b200._native")
    monkeypatch.setattr(N, "require_gpu", lambda: None)
    monkeypatch.setattr(S, "VectorIndex", FakeIndex)
    return S


def _rows(n, start=0):
    return [{"code": f"A{start + i:02d}.{i}", "preferred_zh": f"疾病{start + i}", "has_complication": bool(i % 2), "main_code": f"A{i}",
             "secondary_code": "", "level": 1 + (start + i) % 3, "parent_code": "A", "category_path": f"A > A{i}",
             "semantic_text": f"疾病{start + i} | ICD-10: A{i}"} for i in range(n)]


def _vecs(n, dim=8, seed=0):
    v = np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def _make(S, path, dim=8):
    c = S.IcdStoreClient(uri=path)
    sch = c.create_schema(enable_dynamic_field=True)
    sch.add_field(field_name="id", datatype=S.DataType.INT64, is_primary=True, auto_id=True)
    sch.add_field(field_name="vector", datatype=S.DataType.FLOAT_VECTOR, dim=dim)
    c.create_collection("icd10", schema=sch, index_params=c.prepare_index_params())
    c.load_collection("icd10")
    return c


def test_append_reopen_and_lazy_rows(store, tmp_path):
    S = store
    path = str(tmp_path / "db" / "m.db")
    os.makedirs(os.path.dirname(path))
    c = _make(S, path)
    rows, vecs = _rows(5), _vecs(5)
    rows[2]["note"] = {"dynamic": [1, 2]}                       # enable_dynamic_field: unknown keys are kept
    out = c.insert("icd10", [dict(r, vector=v.tolist()) for r, v in zip(rows, vecs)])
    assert out == {"insert_count": 5, "ids": [0, 1, 2, 3, 4]}
    out = c.insert_arrays("icd10", _rows(3, start=5), _vecs(3, seed=1))
    assert out["ids"] == [5, 6, 7] and c.get_collection_stats("icd10") == {"row_count": 8}
    assert c.row("icd10", 2) == rows[2] and c.field("icd10", "preferred_zh", 6) == "疾病6"
    idx = c.cols["icd10"].index
    assert len(idx) == 8 and idx.levels.tolist() == [1 + i % 3 for i in range(8)]
    c.close()
    # files are flat arrays of exactly the committed size
    d = os.path.join(path + ".icdb", "icd10")
    assert os.path.getsize(os.path.join(d, "vectors.f32")) == 8 * 8 * 4
    assert os.path.getsize(os.path.join(d, "vectors.bf16")) == 8 * 8 * 2
    assert os.path.getsize(os.path.join(d, "levels.u8")) == 8
    assert json.load(open(os.path.join(d, "header.json")))["committed_rows"] == 8
    # a new client maps them back without parsing: same rows, same vectors on the device
    c2 = S.IcdStoreClient(uri=path)
    assert c2.has_collection("icd10") and c2.get_load_state("icd10") == "NotLoad"
    c2.load_collection("icd10")
    got = c2.cols["icd10"].index
    assert np.array_equal(got.vecs[:5], vecs) and got.levels.tolist() == idx.levels.tolist()
    assert c2.row("icd10", 2) == rows[2]
    bf = np.fromfile(os.path.join(d, "vectors.bf16"), "<u2").reshape(8, 8)
    assert np.array_equal(bf[:5], S.f32_to_bf16_bits(vecs))
    c2.release_collection("icd10")
    assert got.closed and c2.get_load_state("icd10") == "NotLoad"
    c2.drop_collection("icd10")
    assert not os.path.exists(d) and not c2.has_collection("icd10")


def test_torn_append_is_discarded_on_open(store, tmp_path):
    S = store
    path = str(tmp_path / "m.db")
    c = _make(S, path)
    c.insert_arrays("icd10", _rows(4), _vecs(4))
    c.close()
    d = os.path.join(path + ".icdb", "icd10")
    # a crash after the column tails were written but before the header was replaced
    with open(os.path.join(d, "vectors.f32"), "ab") as fh:
        fh.write(np.ones(3 * 8, np.float32).tobytes())
    with open(os.path.join(d, "code.str"), "ab") as fh:
        fh.write(b"ORPHAN")
    with open(os.path.join(d, "levels.u8"), "ab") as fh:
        fh.write(b"\x03\x03")
    c2 = S.IcdStoreClient(uri=path)
    assert c2.get_collection_stats("icd10") == {"row_count": 4}
    assert os.path.getsize(os.path.join(d, "vectors.f32")) == 4 * 8 * 4
    c2.load_collection("icd10")
    c2.insert_arrays("icd10", _rows(2, start=4), _vecs(2, seed=5))
    assert [c2.field("icd10", "code", i) for i in range(6)] == [r["code"] for r in _rows(4)] + [r["code"] for r in _rows(2, 4)]
    assert np.array_equal(c2.cols["icd10"].index.vecs[4:], _vecs(2, seed=5))     # row 4 pairs with ITS vector
    # a column file shorter than the committed count is corruption, not something to paper over
    c2.close()
    with open(os.path.join(d, "vectors.f32"), "r+b") as fh:
        fh.truncate(5 * 8 * 4)
    with pytest.raises(Exception, match="committed header needs"):
        S.IcdStoreClient(uri=path)


def test_failed_append_rolls_everything_back(store, tmp_path):
    S = store
    path = str(tmp_path / "m.db")
    c = _make(S, path)
    c.insert_arrays("icd10", _rows(3), _vecs(3))
    d = os.path.join(path + ".icdb", "icd10")
    # (a) a value that cannot be stored is rejected before any file is touched
    bad = _rows(2, start=3)
    bad[1]["level"] = "not-a-number"
    with pytest.raises(Exception):
        c.insert_arrays("icd10", bad, _vecs(2))
    # (b) non-finite vectors and wrong shapes
    with pytest.raises(ValueError):
        c.insert_arrays("icd10", _rows(1, start=3), np.full((1, 8), np.nan, np.float32))
    with pytest.raises(ValueError):
        c.insert_arrays("icd10", _rows(2, start=3), _vecs(1))
    # (c) the device append fails after the column tails were written
    FakeIndex.fail_next = True
    with pytest.raises(RuntimeError, match="injected"):
        c.insert_arrays("icd10", _rows(2, start=3), _vecs(2, seed=9))
    assert c.get_collection_stats("icd10") == {"row_count": 3}
    assert os.path.getsize(os.path.join(d, "vectors.f32")) == 3 * 8 * 4
    assert os.path.getsize(os.path.join(d, "levels.u8")) == 3
    assert json.load(open(os.path.join(d, "header.json")))["committed_rows"] == 3
    assert len(c.cols["icd10"].index) == 3 and c.get_load_state("icd10") == "Loaded"    # rebuilt from the files
    c.insert_arrays("icd10", _rows(2, start=3), _vecs(2, seed=9))
    assert c.get_collection_stats("icd10") == {"row_count": 5} and len(c.cols["icd10"].index) == 5
    c.close()


def test_sharded_clients_load_their_row_ranges(store, tmp_path):
    S = store
    path = str(tmp_path / "m.db")
    c = _make(S, path)
    vecs = _vecs(11, seed=3)
    c.insert_arrays("icd10", _rows(11), vecs)
    c.close()
    seen = []
    for rank in range(3):
        s = S.IcdStoreClient(uri=path, shard=(rank, 3), keep_f32=False)
        s.load_collection("icd10")
        col = s.cols["icd10"]
        lo, hi = S.shard_rows(11, rank, 3)
        assert (col.row_lo, col.row_hi) == (lo, hi) and len(col.index) == hi - lo
        want = (S.f32_to_bf16_bits(vecs[lo:hi]).astype(np.uint32) << 16).view(np.float32)    # the bf16 image was loaded
        assert np.array_equal(col.index.vecs, want) and col.index.levels.tolist() == [1 + i % 3 for i in range(lo, hi)]
        with pytest.raises(Exception, match="read-only"):
            s.insert_arrays("icd10", _rows(1), _vecs(1))
        seen += list(range(lo, hi))
        s.close()
    assert seen == list(range(11))


def test_non_persistent_store_keeps_rows_in_memory(store):
    S = store
    c = S.IcdStoreClient(uri="http://remote:19530")
    sch = c.create_schema()
    sch.add_field(field_name="vector", datatype=S.DataType.FLOAT_VECTOR, dim=8)
    c.create_collection("t", schema=sch)
    c.insert_arrays("t", _rows(4), _vecs(4))
    c.insert_arrays("t", _rows(2, start=4), _vecs(2, seed=2))
    assert c.get_collection_stats("t") == {"row_count": 6} and c.row("t", 5)["code"] == _rows(2, 4)[1]["code"]
    assert len(c.cols["t"].index) == 6
