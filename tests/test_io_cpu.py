"""
Output formats: odil_b200/io.py must write the files the reference's io module writes, byte for
byte, and read them back.  Goldens: tests/golden/io/* written by the unmodified reference
(tests/golden/make_io_goldens.py); the inputs are rebuilt here from the same seeded `cases()`.
"""
import io as pyio
import os

import numpy as np
import pytest

import odil
from odil_b200 import io as oio
from tests.golden.make_io_goldens import cases, write_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")
CASES = cases()


@pytest.mark.parametrize("name", list(CASES))
def test_files_are_byte_identical_to_the_reference(name, tmp_path):
    kind, kw = CASES[name]
    for path in write_case(oio, name, kind, kw, str(tmp_path)):
        with open(path, "rb") as f, open(os.path.join(GOLD, os.path.basename(path)), "rb") as g:
            assert f.read() == g.read(), os.path.basename(path)


@pytest.mark.parametrize("name", [n for n, (k, _) in CASES.items() if k == "xmf"])
def test_read_back_golden_xmf(name):
    _, kw = CASES[name]
    u, meta = oio.read_raw_with_xmf(os.path.join(GOLD, name + ".xmf"))
    ref = np.asarray(kw["u"])
    ref3 = ref if ref.ndim == 3 else ref.reshape((1,) + ref.shape)  # rank 1 stays rank 2, as upstream writes it
    assert u.dtype == ref.dtype and np.array_equal(u, ref3)
    assert meta["count"] == ref3.shape and meta["cell"] == kw["cell"]
    assert meta["name"] == (kw["name"] or "data")
    assert meta["precision"] == ref.dtype.itemsize
    sp = list(kw["spacing"])
    sp = sp if len(sp) == 3 else sp + [min(sp)]
    assert meta["spacing"] == tuple(float(s) for s in sp)
    assert os.path.basename(meta["rawpath"]) == name + ".raw"
    assert oio.read_raw(os.path.join(GOLD, name + ".xmf"))[1] == meta


def test_raw_path_is_relative_to_the_xmf(tmp_path):
    sub = tmp_path / "a"
    sub.mkdir()
    u = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    xmf = oio.write_raw_with_xmf(u, str(sub / "f.xmf"), rawpath=str(tmp_path / "data.raw"), name="f")
    assert "../data.raw" in open(xmf).read()
    back, meta = oio.read_raw_with_xmf(xmf)
    assert np.array_equal(back, u) and os.path.samefile(meta["rawpath"], tmp_path / "data.raw")


def test_unknown_center_is_rejected(tmp_path):
    u = np.zeros((1, 2, 2))
    xmf = oio.write_raw_with_xmf(u, str(tmp_path / "f.xmf"))
    text = open(xmf).read().replace('Center="Cell"', 'Center="Face"')
    open(xmf, "w").write(text)
    with pytest.raises(RuntimeError, match="Unknown Center"):
        oio.parse_raw_xmf(xmf)


def test_vtk_checks_and_file_objects():
    pts = np.zeros((3, 3))
    buf = pyio.BytesIO()
    oio.write_vtk_poly(buf, pts, polygons=[[0, 1, 2]], cell_fields={"a": [1.0]})
    assert buf.getvalue().startswith(b"# vtk DataFile Version 2.0\n\nASCII\nDATASET POLYDATA\nPOINTS 3 float\n")
    assert b"CELL_DATA 1\nSCALARS a float\nLOOKUP_TABLE default\n" in buf.getvalue()
    with pytest.raises(RuntimeError, match="npoints=3"):
        oio.write_vtk_poly(pyio.BytesIO(), pts, point_fields={"a": np.zeros(4)})
    with pytest.raises(RuntimeError, match="ncells=1"):
        oio.write_vtk_poly(pyio.BytesIO(), pts, polygons=[[0, 1, 2]], cell_fields={"a": np.zeros(2)})
    with pytest.raises(RuntimeError, match=r"array.shape=\(3, 2\)"):
        oio.write_vtk_poly(pyio.BytesIO(), pts, tcoords=np.zeros((3, 3)))


def test_public_names_match_the_reference_package():
    # reference src/odil/__init__.py:26-33 exports these at package level
    for name in ["parse_raw_xmf", "read_raw", "read_raw_with_xmf", "write_raw_with_xmf", "write_raw_xmf",
                 "write_vtk_poly"]:
        assert getattr(odil, name) is getattr(oio, name)
    assert odil.io is oio


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_raw_xmf_round_trip(dtype, tmp_path, monkeypatch):
    """The reference's own IO test (tests/test_io.py:12-27) through the package-level names: a linspace field of
    shape (nz, ny, nx) = (5, 4, 3) written next to `data.xdmf2` in the working directory comes back with its shape,
    spacing, name and precision."""
    monkeypatch.chdir(tmp_path)
    nx, ny, nz = 3, 4, 5
    spacing = (4 / nx, 5 / ny, 6 / nz)
    src = np.linspace(0, 1, nx * ny * nz).reshape((nz, ny, nx)).astype(dtype)
    odil.write_raw_with_xmf(src, "data.xdmf2", spacing=spacing, name="foo")
    u, meta = odil.read_raw_with_xmf("data.xdmf2")
    assert meta["count"] == src.shape and u.dtype == dtype and np.array_equal(u, src)
    np.testing.assert_array_almost_equal(meta["spacing"], spacing, decimal=8)
    assert meta["name"] == "foo" and meta["precision"] == np.dtype(dtype).itemsize


def test_history_matches_reference_bytes(tmp_path):
    """train.csv, the saved pickle and the CSV written on reload after the scripted session of
    tests/golden/make_io_goldens.py::history_session are byte-identical to what the reference's History leaves
    (late series, skipped values, array scalars, error messages are asserted inside the session)."""
    from tests.golden.make_io_goldens import history_session

    for path in history_session(odil.History, str(tmp_path)):
        with open(path, "rb") as f, open(os.path.join(GOLD, os.path.basename(path)), "rb") as g:
            assert f.read() == g.read(), os.path.basename(path)


def test_history_takes_device_style_scalars(tmp_path):
    class Lazy:  # stands in for the lazily fetched loss of eval_loss_grad
        def __array__(self, dtype=None, copy=None):
            return np.array(0.25)

    h = odil.History(csvpath=str(tmp_path / "t.csv"))
    h.append("loss", Lazy())
    h.write()
    h.close()
    assert open(tmp_path / "t.csv").read() == "loss\n0.25\n"
    assert h.csvkeys == ["loss"] and h.csvcount == 1 and h.csvpath.endswith("t.csv")


def test_plotutil_savefig_without_matplotlib(tmp_path, monkeypatch):
    """`from odil import plotutil` works on hosts without matplotlib (it is loaded on first use); `savefig` writes one
    file per extension, blanks the time stamps of vector formats, honours ODIL_EXTLIST and skip_existing."""
    from odil import plotutil

    class Fig:
        def __init__(self):
            self.calls = []

        def savefig(self, path, metadata=None, **kw):
            self.calls.append((os.path.basename(path), metadata, kw))
            open(path, "w").close()

    fig, said = Fig(), []
    base = str(tmp_path / "u_00001")
    plotutil.savefig(fig, base, extlist=["png", "pdf", "svg"], printf=said.append, pad_inches=0.01)
    assert fig.calls == [("u_00001.png", {}, {"pad_inches": 0.01}),
                         ("u_00001.pdf", {"CreationDate": None, "DateModified": None}, {"pad_inches": 0.01}),
                         ("u_00001.svg", {"Date": None}, {"pad_inches": 0.01})]
    assert said == [base + ".png", base + ".pdf", base + ".svg"]
    plotutil.savefig(fig, base, extlist=["png"], skip_existing=True, printf=said.append)
    assert len(fig.calls) == 3 and said[-1] == "skip existing '{}.png'".format(base)
    monkeypatch.setenv("ODIL_EXTLIST", "pdf,svg")
    plotutil.set_extlist()
    plotutil.savefig(fig, str(tmp_path / "v"))
    assert [c[0] for c in fig.calls[3:]] == ["v.pdf", "v.svg"]
    monkeypatch.delenv("ODIL_EXTLIST")
    plotutil.set_extlist()
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="needs matplotlib"):
            plotutil.set_log_ticks(None)
