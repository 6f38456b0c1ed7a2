"""Minimal PCD v0.7 reader (ascii / binary / binary_compressed) for x y z rgb[a] clouds.

Host-side I/O only: stands in for pcl::io::loadPCDFile as the reference's CLI uses it
(src/supervoxel_clustering.cpp:313).  Returns pcl::PointXYZRGBA-layout points.
"""
import struct
import numpy as np
from .synth import POINT_DTYPE


def _lzf_decompress(data, out_len):
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(data)
    while ip < n:
        ctrl = data[ip]
        ip += 1
        if ctrl < 32:
            ln = ctrl + 1
            out[op:op + ln] = data[ip:ip + ln]
            ip += ln
            op += ln
        else:
            ln = ctrl >> 5
            if ln == 7:
                ln += data[ip]
                ip += 1
            ref = op - ((ctrl & 0x1F) << 8) - data[ip] - 1
            ip += 1
            ln += 2
            if ref + ln <= op:
                out[op:op + ln] = out[ref:ref + ln]
                op += ln
            else:                      # overlapping back-reference: byte-wise copy
                for _ in range(ln):
                    out[op] = out[ref]
                    op += 1
                    ref += 1
    return bytes(out[:op])


def read_pcd(path):
    with open(path, "rb") as f:
        raw = f.read()
    hdr = {}
    pos = 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        k, _, v = line.partition(" ")
        hdr[k.upper()] = v.split()
        if k.upper() == "DATA":
            break
    fields = hdr["FIELDS"]
    sizes = [int(s) for s in hdr["SIZE"]]
    types = hdr["TYPE"]
    counts = [int(c) for c in hdr.get("COUNT", ["1"] * len(fields))]
    n = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr["HEIGHT"][0])
    np_t = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4",
            ("I", 1): "i1", ("I", 2): "<i2", ("I", 4): "<i4"}
    dts = [np_t[(t, s)] for t, s in zip(types, sizes)]
    mode = hdr["DATA"][0]
    cols = {}
    if mode == "ascii":
        arr = np.loadtxt(raw[pos:].decode().splitlines(), ndmin=2)
        for i, name in enumerate(fields):
            cols[name] = arr[:, i].astype(dts[i])
    elif mode == "binary":
        rec = np.dtype([(nm, dt, (c,)) if c > 1 else (nm, dt) for nm, dt, c in zip(fields, dts, counts)])
        arr = np.frombuffer(raw, dtype=rec, count=n, offset=pos)
        for name in fields:
            cols[name] = arr[name]
    elif mode == "binary_compressed":
        csz, usz = struct.unpack_from("<II", raw, pos)
        data = _lzf_decompress(raw[pos + 8:pos + 8 + csz], usz)
        off = 0
        for name, dt, c in zip(fields, dts, counts):
            nb = np.dtype(dt).itemsize * c * n
            cols[name] = np.frombuffer(data, dtype=dt, count=n * c, offset=off)
            off += nb
    else:
        raise ValueError("unsupported PCD DATA mode " + mode)
    pts = np.zeros(n, dtype=POINT_DTYPE)
    pts["x"], pts["y"], pts["z"] = cols["x"], cols["y"], cols["z"]
    pts["w"] = 1.0
    cname = "rgba" if "rgba" in cols else ("rgb" if "rgb" in cols else None)
    if cname:
        c = cols[cname]
        pts["rgba"] = c.view(np.uint32) if c.dtype.itemsize == 4 else c.astype(np.uint32)
    label = cols.get("label")
    return pts, (label.astype(np.uint32) if label is not None else None), hdr


def write_pcd_binary(path, pts):
    """pcl::PointXYZRGBA-layout points -> PCD v0.7 DATA binary with FIELDS x y z rgba."""
    n = len(pts)
    rec = np.zeros(n, dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")]))
    for k in ("x", "y", "z", "rgba"):
        rec[k] = pts[k]
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgba\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n"
           "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(rec.tobytes())


def _lzf_compress_literal(data):
    """A valid LZF stream that uses literal runs only (no back-references): what a decoder must accept, no more."""
    out = bytearray()
    for i in range(0, len(data), 32):
        chunk = data[i:i + 32]
        out.append(len(chunk) - 1)
        out += chunk
    return bytes(out)


def write_pcd(path, pts, label=None, mode="binary", rgb_as_float=False):
    """Test / tooling writer: FIELDS x y z rgba|rgb [label] in DATA ascii | binary | binary_compressed.
    rgb_as_float: declare the colour field as `rgb`, TYPE F (PCL's PointXYZRGB); in ascii it is then printed the way
    PCL >= 1.8 prints it, as the uint32 reinterpretation of the packed colour."""
    n = len(pts)
    cname, ctype = ("rgb", "F") if rgb_as_float else ("rgba", "U")
    names = ["x", "y", "z", cname] + (["label"] if label is not None else [])
    types = ["F", "F", "F", ctype] + (["U"] if label is not None else [])
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\n"
           "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (
               " ".join(names), " ".join(["4"] * len(names)), " ".join(types), " ".join(["1"] * len(names)), n, n, mode))
    cols = [np.ascontiguousarray(pts[k], "<f4") for k in ("x", "y", "z")] + [np.ascontiguousarray(pts["rgba"], "<u4")]
    if label is not None:
        cols.append(np.ascontiguousarray(label, "<u4"))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        if mode == "ascii":
            lines = []
            for i in range(n):
                t = ["nan" if np.isnan(c[i]) else repr(float(c[i])) for c in cols[:3]] + [str(int(c[i])) for c in cols[3:]]
                lines.append(" ".join(t))
            f.write(("\n".join(lines) + "\n").encode("ascii"))
        elif mode == "binary":
            rec = np.zeros(n, dtype=np.dtype([(nm, "<u4") for nm in names]))
            for nm, c in zip(names, cols):
                rec[nm] = c.view("<u4")
            f.write(rec.tobytes())
        elif mode == "binary_compressed":
            soa = b"".join(c.tobytes() for c in cols)
            comp = _lzf_compress_literal(soa)
            f.write(struct.pack("<II", len(comp), len(soa)))
            f.write(comp)
        else:
            raise ValueError(mode)
