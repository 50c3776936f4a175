"""
Field dumps in the formats the reference writes, so downstream viewers and the examples' plot scripts keep working
(reference src/odil/io.py: `write_raw_xmf` :61-144, `write_raw_with_xmf` :147-171, `parse_raw_xmf` :7-43,
`read_raw_with_xmf` :46-58, `write_vtk_poly` :174-280).  Host side only: device fields are brought to the host by
the caller (`domain.field(state, key)` / `np.array(...)`) before they get here.

  RAW + XMF   a flat binary array (C order, (Nz, Ny, Nx)) next to an XDMF-2 "3DCORECTMesh" description that points
              at it through a HyperSlab item; arrays of lower rank get leading unit axes, spacing is stored (x, y, z).
  legacy VTK  POLYDATA with points, polygons and/or lines, scalar point / cell fields, texture coordinates;
              ASCII or big-endian binary.

The files are byte-identical to the reference's (tests/test_io_cpu.py compares against files written by the
reference's own module, tests/golden/make_io_goldens.py).
"""
import os
import re

import numpy as np

_XMF_HEAD = ['<?xml version="1.0" ?>', '<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>', '<Xdmf Version="2.0">']


def _join(values):
    return " ".join(str(v) for v in values)


def _xmf_lines(rawpath, count, spacing, name, precision, cell):
    """The XDMF description, one (indent, text) pair per line."""
    dim = 3
    zyx = _join(count)
    nodes = _join([c + 1 for c in count] if cell else list(count))
    vec = 'Dimensions="{}" NumberType="Float" Precision="8" Format="XML"'.format(dim)
    number = "Double" if precision == 8 else "Float"
    return [
        (1, "<Domain>"),
        (3, '<Grid Name="mesh" GridType="Uniform">'),
        (5, '<Topology TopologyType="{}DCORECTMesh" Dimensions="{}"/>'.format(dim, nodes)),
        (5, '<Geometry GeometryType="ORIGIN_DXDYDZ">'),
        (7, '<DataItem Name="Origin" {}>'.format(vec)),
        (9, _join([0] * dim)),
        (7, "</DataItem>"),
        (7, '<DataItem Name="Spacing" {}>'.format(vec)),
        (9, _join(reversed(list(spacing)))),
        (7, "</DataItem>"),
        (5, "</Geometry>"),
        (5, '<Attribute Name="{}" AttributeType="Scalar" Center="{}">'.format(name, "Cell" if cell else "Node")),
        (7, '<DataItem ItemType="HyperSlab" Dimensions="{}" Type="HyperSlab">'.format(zyx)),
        (11, '<DataItem Dimensions="3 {}" Format="XML">'.format(dim)),
        (13, _join([0] * dim)),  # start
        (13, _join([1] * dim)),  # stride
        (13, zyx),               # count
        (11, "</DataItem>"),
        (11, '<DataItem Dimensions="{}" Seek="0" Precision="{}" NumberType="{}" Format="Binary">'.format(
            zyx, precision, number)),
        (13, str(rawpath)),
        (11, "</DataItem>"),
        (7, "</DataItem>"),
        (5, "</Attribute>"),
        (3, "</Grid>"),
        (1, "</Domain>"),
    ]


def write_raw_xmf(xmfpath, rawpath, count, spacing=(1, 1, 1), name=None, precision=8, cell=True):
    """
    Writes the XMF description of a RAW file.
    xmfpath: output `.xmf`;  rawpath: the binary file as it should be referenced from the XMF;
    count: array shape (Nz, Ny, Nx);  spacing: (hx, hy, hz);  precision: 4 or 8 bytes per value;
    cell: values at cell centres (else at nodes).
    """
    name = "data" if name is None else name
    body = [" " * ind + text for ind, text in _xmf_lines(rawpath, tuple(count), spacing, name, precision, cell)]
    with open(xmfpath, "w") as f:
        f.write("\n".join(_XMF_HEAD + body + ["</Xdmf>"]) + "\n")


def write_raw_with_xmf(u, xmfpath, rawpath=None, spacing=(1, 1, 1), cell=True, name=None):
    """
    Writes `u` as RAW binary plus its XMF description and returns `xmfpath`.
    u: array of rank <= 3, stored as (Nz, Ny, Nx) (lower ranks get a leading unit axis);
    rawpath defaults to `xmfpath` with the extension replaced by `.raw`; float32 stays float32, anything else is
    described as 8-byte values (the array is written as it is).
    """
    u = np.asarray(u)
    if u.ndim != 3:
        u = u.reshape((1,) + u.shape)
    spacing = list(spacing)
    if len(spacing) != 3:
        spacing = spacing + [min(spacing)]
    if rawpath is None:
        rawpath = os.path.splitext(xmfpath)[0] + ".raw"
    link = os.path.relpath(rawpath, start=os.path.dirname(xmfpath))
    write_raw_xmf(xmfpath, link, u.shape, spacing, name, 4 if u.dtype == np.float32 else 8, cell)
    u.tofile(rawpath)
    return xmfpath


_RE_BINARY = re.compile(r'<DataItem[^>]*Dimensions="([\d ]*)"[^>]*Precision="(\d*)"[^>]*Format="Binary"> *([^ <]*)')
_RE_ATTRIBUTE = re.compile(r'<Attribute Name="([^"]*)" AttributeType="Scalar" Center="([a-zA-Z]*)">')
_RE_SPACING = re.compile(r'<DataItem Name="Spacing".*?> *(.*?)<')


def parse_raw_xmf(xmfpath):
    """Metadata of an XMF file written by `write_raw_xmf`: dict with rawpath (resolved against the XMF's folder),
    count (Nz, Ny, Nx), spacing (hx, hy, hz), name, precision, cell.  Reads everything the reference's parser reads;
    in addition RAW links with folders in them (`../data.raw`) and the two-number `Dimensions` that a rank-1 array
    gets from `write_raw_with_xmf` are understood (the reference's pattern, io.py:15-20, stops at both)."""
    with open(xmfpath) as f:
        text = "".join(f.read().split("\n"))
    dims, precision, raw = _RE_BINARY.findall(text)[0]
    name, center = _RE_ATTRIBUTE.findall(text)[0]
    if center not in ("Cell", "Node"):
        raise RuntimeError("Unknown Center='{}'".format(center))
    spacing = tuple(float(v) for v in reversed(_RE_SPACING.findall(text)[0].split()))
    return {
        "rawpath": os.path.join(os.path.dirname(xmfpath), raw),
        "count": tuple(int(v) for v in dims.split()),
        "spacing": spacing,
        "name": name,
        "precision": int(precision),
        "cell": center == "Cell",
    }


def read_raw_with_xmf(xmfpath):
    """Returns (array of shape count, metadata) of a scalar field stored as RAW + XMF."""
    meta = parse_raw_xmf(xmfpath)
    dtype = {4: np.float32, 8: np.float64}[meta["precision"]]
    return np.fromfile(meta["rawpath"], dtype).reshape(meta["count"]), meta


def read_raw(xmfpath):
    return read_raw_with_xmf(xmfpath)


def write_vtk_poly(fout, points, polygons=None, lines=None, point_fields=None, cell_fields=None, tcoords=None,
                   comment="", fmt="%.16g", binary=False):
    """
    Writes a legacy-VTK POLYDATA file.
    fout: path or binary file object;  points: (npoints, 3);  polygons / lines: lists of index lists into `points`;
    point_fields / cell_fields: name -> scalar array over points / polygons;  tcoords: (npoints, 2) texture
    coordinates;  binary: big-endian float32 payload instead of text formatted with `fmt`.
    """
    own = isinstance(fout, str)
    f = open(fout, "wb") if own else fout
    try:
        def line(text=""):
            f.write((text if isinstance(text, bytes) else text.encode()) + b"\n")

        def floats(a):
            if binary:
                np.asarray(a, dtype=">f").tofile(f)
            else:
                np.savetxt(f, a, fmt=fmt)

        def connectivity(keyword, cells, as_binary):
            line("{} {} {}".format(keyword, len(cells), len(cells) + sum(len(c) for c in cells)))
            for c in cells:
                if as_binary:
                    np.array([len(c)] + list(c), dtype=">i4").tofile(f)
                else:
                    line(_join([len(c)] + list(c)))

        def scalars(fields, expected, what):
            for key, a in fields.items():
                a = np.reshape(a, -1)
                if a.size != expected:
                    raise RuntimeError(f"Expected equal array.size={a.size} and {what}={expected}")
                line("SCALARS {} float".format(key))
                line("LOOKUP_TABLE default")
                floats(a)

        npoints = len(points)
        line("# vtk DataFile Version 2.0")
        line(comment)
        line("BINARY" if binary else "ASCII")
        line("DATASET POLYDATA")
        line("POINTS {} float".format(npoints))
        floats(points)
        ncells = None
        if polygons is not None:
            ncells = len(polygons)
            connectivity("POLYGONS", polygons, False)  # polygon connectivity is text in both modes, as upstream
        if lines is not None:
            connectivity("LINES", lines, binary)
        if point_fields is not None or tcoords is not None:
            line("POINT_DATA {}".format(npoints))
        if point_fields is not None:
            scalars(point_fields, npoints, "npoints")
        if tcoords is not None:
            if np.shape(tcoords) != (npoints, 2):
                raise RuntimeError("Expected array.shape=({}, 2), got {}".format(npoints, np.shape(tcoords)))
            line("TEXTURE_COORDINATES tcoords 2 float")
            floats(tcoords)
        if cell_fields is not None:
            if ncells is None:
                raise RuntimeError("cell_fields need polygons")
            line("CELL_DATA {}".format(ncells))
            scalars(cell_fields, ncells, "ncells")
    finally:
        if own:
            f.close()
