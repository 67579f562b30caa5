"""TF V2 checkpoint reader/writer (pfnl_b200/tf_checkpoint.py) - CPU only.
The reference restores with tf.train.Saver (base_model.py:231-243); no TF-written file exists offline,
so the format is pinned by CRC-32C known answers, hand-assembled LevelDB blocks and round trips."""
import os
import struct

import numpy as np
import pytest

from pfnl_b200 import tf_checkpoint as T
from pfnl_b200 import weights as WT


def test_crc32c_known_answers(built_lib):
    # RFC 3720 B.4 vectors + the classic check value
    for fn in (T.crc32c_py, T.crc32c):
        assert fn(b"123456789") == 0xE3069283
        assert fn(bytes(32)) == 0x8A9136AA
        assert fn(b"\xff" * 32) == 0x62A8AB43
        assert fn(bytes(range(32))) == 0x46DD794E
        assert fn(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert T.crc32c_py(b"") == 0


def test_native_crc_matches_python_on_large_and_ragged_buffers(built_lib):
    from pfnl_b200._lib import lib
    rng = np.random.default_rng(5)
    for n in (1, 7, 8, 9, 4095, 4096, 4097, 100003):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert int(lib.pfnl_crc32c(b, n, 0)) == T.crc32c_py(b)
    # continuation: crc(a+b) == crc(b, crc(a))
    a, b = b"hello, ", b"tensor bundle"
    assert int(lib.pfnl_crc32c(b, len(b), T.crc32c_py(a))) == T.crc32c_py(a + b)
    assert T.crc32c_py(b, T.crc32c_py(a)) == T.crc32c_py(a + b)


def test_crc_mask_roundtrip_and_leveldb_property():
    c = T.crc32c_py(b"foo")
    assert T.mask_crc(c) != c and T.mask_crc(T.mask_crc(c)) != c
    assert T.unmask_crc(T.mask_crc(c)) == c
    assert T.unmask_crc(T.unmask_crc(T.mask_crc(T.mask_crc(c)))) == c


def test_hand_assembled_block_with_prefix_compression():
    # leveldb table_format: [shared][non_shared][value_len][key delta][value] ... restarts[] num_restarts
    blk = bytes([0, 5, 1]) + b"apple" + b"1"
    blk += bytes([2, 5, 2]) + b"ricot" + b"22"      # "ap" shared -> "apricot"
    blk += bytes([0, 6, 0]) + b"banana"             # restart point, empty value
    blk += struct.pack("<III", 0, 19, 2)
    assert list(T._block_entries(blk)) == [(b"apple", b"1"), (b"apricot", b"22"), (b"banana", b"")]


def test_table_roundtrip_multiblock_and_corruption(tmp_path):
    items = [(f"key/{i:05d}".encode(), os.urandom(i % 50)) for i in range(1000)]
    p = str(tmp_path / "t.index")
    T.write_table(p, items, block_size=512)
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == T.TABLE_MAGIC and len(raw) > 48
    assert T.read_table(p) == sorted(items)
    bad = bytearray(raw)
    bad[100] ^= 0x40
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        T.read_table(p)
    open(p, "wb").write(raw[:-1] + b"\x00")
    with pytest.raises(ValueError, match="magic"):
        T.read_table(p)


def test_empty_table(tmp_path):
    p = str(tmp_path / "e.index")
    T.write_table(p, [])
    assert T.read_table(p) == []


def _ckpt(tmp_path, step=150000):
    w = WT.xavier_init(seed=7)
    extra = dict(w)
    extra["global_step"] = np.array(step, np.int64)
    extra["beta1_power"] = np.array(0.5, np.float32)
    extra["nlvsr/conv0/kernel/Adam"] = np.ones((5, 5, 3, 64), np.float32)      # optimizer slots are ignored on load
    extra["nlvsr/conv0/kernel/Adam_1"] = np.full((5, 5, 3, 64), 2, np.float32)
    d = tmp_path / "ckpt"
    d.mkdir()
    T.write_bundle(str(d / f"VSR-{step}"), extra)
    return str(d), w, extra


def test_bundle_roundtrip_all_dtypes_and_subset(tmp_path):
    d, w, extra = _ckpt(tmp_path)
    prefix = os.path.join(d, "VSR-150000")
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(a.nbytes for a in map(np.asarray, extra.values()))
    lv = T.list_variables(prefix)
    assert set(lv) == set(extra)
    assert lv["nlvsr/convmerge1/kernel"] == (np.float32, (3, 3, 448, 48)) and lv["global_step"] == (np.int64, ())
    got = T.read_bundle(prefix)
    for k, v in extra.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape
        assert np.array_equal(got[k], v), k
    sub = T.read_bundle(prefix, ["nlvsr/conv1_3/bias", "global_step"])
    assert set(sub) == {"nlvsr/conv1_3/bias", "global_step"} and int(sub["global_step"]) == 150000
    with pytest.raises(KeyError):
        T.read_bundle(prefix, ["nlvsr/conv1_20/kernel"])


def test_bundle_detects_corrupt_tensor_bytes(tmp_path):
    d, _, _ = _ckpt(tmp_path)
    data = os.path.join(d, "VSR-150000.data-00000-of-00001")
    raw = bytearray(open(data, "rb").read())
    raw[len(raw) // 2] ^= 1
    open(data, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        T.read_bundle(os.path.join(d, "VSR-150000"))
    T.read_bundle(os.path.join(d, "VSR-150000"), verify_crc=False)


def test_checkpoint_state_file(tmp_path):
    d, _, _ = _ckpt(tmp_path)
    assert T.read_checkpoint_state(d) is None and T.latest_checkpoint(d) is None
    T.write_checkpoint_state(d, "VSR-150000", keep=["VSR-149500"])
    assert T.read_checkpoint_state(d) == "VSR-150000"
    assert T.latest_checkpoint(d) == os.path.join(d, "VSR-150000")
    # the reference keeps only the basename (base_model.py:237): an absolute path from another machine works
    open(os.path.join(d, "checkpoint"), "wt").write(
        'model_checkpoint_path: "/home/someone/checkpoint/pfnl/VSR-150000"\n'
        'all_model_checkpoint_paths: "/home/someone/checkpoint/pfnl/VSR-150000"\n')
    assert T.latest_checkpoint(d) == os.path.join(d, "VSR-150000")
    open(os.path.join(d, "checkpoint"), "wt").write('model_checkpoint_path: "VSR-1"\n')
    assert T.latest_checkpoint(d) is None


def test_pfnl_load_and_save_like_base_model(tmp_path, capsys, built_lib):
    from pfnl_b200 import PFNL
    d, w, _ = _ckpt(tmp_path, step=4500)
    m = PFNL()
    assert m.load(str(tmp_path / "nothing_here")) is False           # base_model.py:241-243
    assert "ERROR" in capsys.readouterr().out
    T.write_checkpoint_state(d, "VSR-4500")
    assert m.load(d) is True
    assert "VSR-4500 Success" in capsys.readouterr().out
    assert m.global_step == 4500
    for k, v in w.items():
        assert np.array_equal(m._weights[k], v)
    # save -> a fresh instance restores the same weights; the state file names the new checkpoint
    m.global_step = 5000
    out = m.save(str(tmp_path / "resaved"))
    assert out.endswith("VSR-5000") and T.read_checkpoint_state(str(tmp_path / "resaved")) == "VSR-5000"
    m2 = PFNL()
    assert m2.load(str(tmp_path / "resaved")) and m2.global_step == 5000
    for k, v in w.items():
        assert np.array_equal(m2._weights[k], v)
    # the re-saved checkpoint names the step like the reference's unnamed counter does (pfnl.py:152): 'Variable', int32
    lv = T.list_variables(out)
    assert lv["Variable"] == (np.int32, ()) and "global_step" not in lv
    # a reference-written checkpoint with neither key: the step comes from the VSR-<step> file name
    (tmp_path / "nostep").mkdir()
    T.write_bundle(str(tmp_path / "nostep" / "VSR-777"), w)
    T.write_checkpoint_state(str(tmp_path / "nostep"), "VSR-777")
    m3 = PFNL()
    assert m3.load(str(tmp_path / "nostep")) and m3.global_step == 777
    # a checkpoint that lacks a model variable is an error, not a silent partial restore
    bad = {k: v for k, v in w.items() if k != "nlvsr/conv2_7/bias"}
    (tmp_path / "bad").mkdir()
    T.write_bundle(str(tmp_path / "bad" / "VSR-1"), bad)
    T.write_checkpoint_state(str(tmp_path / "bad"), "VSR-1")
    with pytest.raises(KeyError):
        PFNL().load(str(tmp_path / "bad"))


def test_varint_and_table_roundtrip_property():
    """Random keys / values through the SSTable writer and reader (hypothesis): sorted output, every pair
    intact, for block sizes that force one entry per block as well as one block in total."""
    from hypothesis import given, settings, strategies as st
    import tempfile

    @settings(max_examples=40, deadline=None)
    @given(st.dictionaries(st.binary(min_size=0, max_size=40), st.binary(min_size=0, max_size=300), max_size=60),
           st.sampled_from([1, 64, 4096, 1 << 20]))
    def run(d, block_size):
        items = list(d.items())
        with tempfile.TemporaryDirectory() as tmp:
            p = os.path.join(tmp, "t.index")
            T.write_table(p, items, block_size=block_size)
            assert T.read_table(p) == sorted(items)

    run()

    @settings(max_examples=200, deadline=None)
    @given(st.integers(min_value=0, max_value=(1 << 64) - 1))
    def varint(v):
        b = T._put_varint(v)
        assert T._get_varint(b + b"\xff", 0) == (v, len(b)) and len(b) <= 10

    varint()


def test_bundle_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.index"
    p.write_bytes(b"not a table")
    with pytest.raises(ValueError):
        T.read_table(str(p))
    p.write_bytes(bytes(100))
    with pytest.raises(ValueError, match="magic"):
        T.read_table(str(p))
