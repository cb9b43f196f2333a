"""the sparse-GH table file in the reference's wire format (cereal BinaryOutputArchive of
unordered_map<tuple<double,double>, tuple<MatrixXd,VectorXd>>, quadrature/saveSparseGHWeightMap.h:43-51,
helpers/SerializeEigenMaps.h:195-224).

Golden file: tests/golden/spgh_cereal/spgh_small_cereal.bin was written by the REAL cereal library (headers vendored in
the reference tree) through tests/golden/spgh_cereal/cereal_table_tool.cpp; the tool is rebuilt and used as an
independent reader of OUR files whenever the reference tree is present (this container), otherwise only the
committed golden bytes are used.  No GPU needed: the file I/O is host code of libgvib200.so."""
import pathlib
import struct
import subprocess

import numpy as np
import pytest

import gaussianvi_b200 as gv
from gaussianvi_b200 import capi
from oracle import gvi_oracle

HERE = pathlib.Path(__file__).resolve().parent
GOLD = HERE / "golden" / "spgh_cereal" / "spgh_small_cereal.bin"
TOOL_SRC = HERE / "golden" / "spgh_cereal" / "cereal_table_tool.cpp"
CEREAL_INC = pathlib.Path("/root/reference/include/cereal/include")


def parse(path):
    """Independent pure-Python parser of the wire format -> {(dim, deg): (nodes[n, dim], w[n])} and raw entry bytes."""
    b = pathlib.Path(path).read_bytes()
    (count,), off = struct.unpack_from("<Q", b, 0), 8
    out, raw = {}, {}
    for _ in range(count):
        start = off
        dim, deg = struct.unpack_from("<dd", b, off); off += 16
        rows, cols = struct.unpack_from("<ii", b, off); off += 8
        nodes = np.frombuffer(b, dtype="<f8", count=rows * cols, offset=off).reshape(rows, cols); off += 8 * rows * cols
        (size,) = struct.unpack_from("<i", b, off); off += 4
        w = np.frombuffer(b, dtype="<f8", count=size, offset=off); off += 8 * size
        out[(int(dim), int(deg))] = (nodes, w)
        raw[(int(dim), int(deg))] = b[start:off]
    assert off == len(b)
    return out, raw


def test_golden_file_matches_oracle_tables():
    got, _ = parse(GOLD)
    assert sorted(got) == [(1, 3), (2, 2), (3, 2)]
    for (dim, deg), (nodes, w) in got.items():
        Z, W = gvi_oracle.table(dim, deg)
        np.testing.assert_array_equal(nodes, Z)
        np.testing.assert_array_equal(w, W)


def test_query_reads_the_cereal_written_file():
    assert sorted(capi.table_file_query(GOLD)) == [(1, 3, 3), (2, 2, 5), (3, 2, 7)]


def test_writer_is_byte_identical_per_entry(tmp_path):
    out = tmp_path / "ours.bin"
    capi.table_file_write(out, [(1, 3), (2, 2), (3, 2)])
    ours, raw_ours = parse(out)
    _, raw_gold = parse(GOLD)
    assert raw_ours == raw_gold                      # every entry record, byte for byte
    assert out.stat().st_size == GOLD.stat().st_size  # same framing (uint64 count + the records)


def test_roundtrip_headline_rules(tmp_path):
    keys = [(4, 6), (12, 4), (1, 10)]
    out = tmp_path / "t.bin"
    capi.table_file_write(out, keys)
    got, _ = parse(out)
    for dim, deg in keys:
        Z, W = gvi_oracle.table(dim, deg)
        np.testing.assert_array_equal(got[(dim, deg)][0], Z)
        np.testing.assert_array_equal(got[(dim, deg)][1], W)
    assert [(d, k) for d, k, _ in capi.table_file_query(out)] == keys


def test_malformed_files_are_rejected(tmp_path):
    bad = tmp_path / "bad.bin"
    bad.write_bytes(GOLD.read_bytes()[:-5])
    with pytest.raises(capi.GviError):
        capi.table_file_query(bad)
    bad.write_bytes(GOLD.read_bytes() + b"\0")
    with pytest.raises(capi.GviError):
        capi.table_file_query(bad)
    with pytest.raises(capi.GviError):
        capi.table_file_query(tmp_path / "missing.bin")
    with pytest.raises(capi.GviError):
        capi.table_file_write(tmp_path / "x.bin", [(4, 99)])   # no such rule


@pytest.mark.skipif(not CEREAL_INC.exists(), reason="reference tree (vendored cereal headers) not present")
def test_real_cereal_reads_our_file(tmp_path):
    tool = tmp_path / "cereal_table_tool"
    subprocess.run(["g++", "-std=c++17", "-O1", f"-I{CEREAL_INC}", f"-I{HERE.parent / 'gaussianvi_b200' / 'csrc'}",
                    str(TOOL_SRC), str(HERE.parent / "gaussianvi_b200" / "csrc" / "spgh_table.cpp"), "-o", str(tool)],
                   check=True)
    out = tmp_path / "ours.bin"
    keys = [(2, 3), (4, 6), (5, 2)]
    capi.table_file_write(out, keys)
    txt = subprocess.run([str(tool), "read", str(out)], check=True, capture_output=True, text=True).stdout.split("\n")
    i = 0
    for dim, deg in sorted(keys):
        Z, W = gvi_oracle.table(dim, deg)
        assert [int(v) for v in txt[i].split()] == [dim, deg, len(W), dim]
        rows = np.array([[float(v) for v in txt[i + 1 + r].split()] for r in range(len(W))])
        np.testing.assert_array_equal(rows[:, :dim], Z)
        np.testing.assert_array_equal(rows[:, dim], W)
        i += 1 + len(W)
    # and the golden file regenerates bit-identically apart from the unordered_map's entry order
    regen = tmp_path / "regen.bin"
    subprocess.run([str(tool), "write", str(regen)], check=True)
    assert parse(regen)[1] == parse(GOLD)[1]


@pytest.mark.gpu
def test_loaded_table_drives_the_device_path(tmp_path):
    """A context that loads a file holds exactly the file's rules."""
    out = tmp_path / "t.bin"
    capi.table_file_write(out, [(1, 10)])
    ctx = gv.Context(0)
    assert ctx.table_file_load(out) == 1
    Z, W = ctx.table_get(1, 10)
    Zo, Wo = gvi_oracle.table(1, 10)
    np.testing.assert_array_equal(Z, Zo)
    np.testing.assert_array_equal(W, Wo)
