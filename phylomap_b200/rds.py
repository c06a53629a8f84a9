"""Reader and writer for R's serialisation format (readRDS / saveRDS), the container of the reference's fixtures:
`inst/extdata/Squamate/phylomap_compatible_squamate_tree.RData` is written with saveRDS (R/Squamate_tree_setup.R:85) and
read with readRDS (vignettes/Squamate_DIC_model_selection.Rnw:78); the DIC helper saves every trace with saveRDS
(R/sourceme.R:533-595).  SURVEY.md §8(f) rank 4.

Only what those objects need: XDR ("X\\n") streams of version 2 or 3, gzip / bzip2 / xz / no compression; NULL, logical,
integer, double, character and generic vectors (lists), pairlists (attributes), symbols and their back references.
Environments, closures, S4 objects and ALTREP payloads are refused with a clear error.

    obj = read_rds(path)     # list with names -> dict (insertion order = R order); vectors -> numpy arrays; a `dim`
                             # attribute reshapes column-major; other attributes are kept in RObject.attributes
    write_rds(path, obj)     # dict -> named list, list/tuple -> list, numpy arrays -> vectors / matrices (+ dimnames)
"""
import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE_SXP, SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 254, 1, 2, 9, 10, 13, 14, 16, 19
REFSXP, NAMESPACESXP, ALTREP_SXP, ATTRLISTSXP, ATTRLANGSXP = 255, 249, 238, 239, 240
NA_INTEGER = -2147483648


class RObject:
    """A value that carries R attributes other than names / dim (class, levels, dimnames, ...)."""

    def __init__(self, value, attributes):
        self.value = value
        self.attributes = attributes

    def __repr__(self):
        return "RObject(%r, attributes=%r)" % (self.value, list(self.attributes))


def _decompress(raw):
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


class _Reader:
    def __init__(self, data):
        self.d = data
        self.p = 0
        self.refs = []

    def take(self, n):
        b = self.d[self.p:self.p + n]
        if len(b) != n:
            raise ValueError("truncated RDS stream")
        self.p += n
        return b

    def int(self):
        return struct.unpack(">i", self.take(4))[0]

    def length(self):
        n = self.int()
        if n == -1:  # long vector: two more words
            hi, lo = struct.unpack(">II", self.take(8))
            return (hi << 32) | lo
        return n

    def item(self):
        flags = self.int()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == NILVALUE_SXP:
            return None
        if typ == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if typ == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if typ in (LISTSXP, ATTRLISTSXP):  # pairlist: returned as an ordered list of (tag, value)
            out = []
            while True:
                attrs = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                del attrs
                flags = self.int()
                typ = flags & 0xFF
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
                if typ == NILVALUE_SXP:
                    return out
                if typ not in (LISTSXP, ATTRLISTSXP):
                    raise ValueError("malformed pairlist in RDS stream")
        if typ == CHARSXP:
            n = self.int()
            if n == -1:
                return None  # NA_character_
            return self.take(n).decode("utf-8", errors="replace")
        if typ == LGLSXP:
            n = self.length()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
            val = v.astype(bool) if not np.any(v == NA_INTEGER) else np.where(v == NA_INTEGER, np.nan, v).astype(float)
        elif typ == INTSXP:
            n = self.length()
            val = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif typ == REALSXP:
            n = self.length()
            val = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif typ == STRSXP:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif typ == VECSXP:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif typ == ALTREP_SXP:
            raise ValueError("ALTREP objects are not supported (save with saveRDS(..., version = 2))")
        else:
            raise ValueError("unsupported R type %d in RDS stream" % typ)
        if has_attr:
            attrs = dict((k, v) for k, v in self.item())
            return _apply_attributes(val, attrs)
        return val


def _apply_attributes(val, attrs):
    names = attrs.pop("names", None)
    dim = attrs.pop("dim", None)
    if dim is not None and isinstance(val, np.ndarray):
        val = val.reshape(tuple(int(x) for x in dim), order="F")
    if names is not None and isinstance(val, list):
        if all(n is not None and n != "" for n in names) and len(set(names)) == len(names):
            val = dict(zip(names, val))
        else:
            attrs["names"] = names
    elif names is not None:
        attrs["names"] = names
    return RObject(val, attrs) if attrs else val


def read_rds(path):
    """readRDS(path)."""
    with open(path, "rb") as f:
        data = _decompress(f.read())
    if data[:2] != b"X\n":
        raise ValueError("only XDR-format RDS files are supported (saveRDS default)")
    r = _Reader(data)
    r.p = 2
    version = r.int()
    r.int()  # R version that wrote the file
    r.int()  # minimal R version to read it
    if version == 3:
        r.take(r.int())  # native encoding
    elif version != 2:
        raise ValueError("unsupported RDS version %d" % version)
    return r.item()


def plain(obj):
    """Strip RObject wrappers (drop class / dimnames ...) recursively."""
    if isinstance(obj, RObject):
        return plain(obj.value)
    if isinstance(obj, dict):
        return {k: plain(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [plain(v) for v in obj]
    return obj


# ---- writer ------------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.out = []
        self.syms = {}

    def int(self, v):
        self.out.append(struct.pack(">i", v))

    def charsxp(self, s):
        if s is None:
            self.int(CHARSXP)
            self.int(-1)
            return
        b = s.encode("utf-8")
        self.int(CHARSXP | 0x8000)  # UTF-8 flag in the gp field (bit 3 of gp << 12)
        self.int(len(b))
        self.out.append(b)

    def symbol(self, name):
        if name in self.syms:
            self.int(REFSXP | (self.syms[name] << 8))
            return
        self.syms[name] = len(self.syms) + 1
        self.int(SYMSXP)
        self.charsxp(name)

    def attributes(self, attrs):
        for k, v in attrs.items():
            self.int(LISTSXP | 0x400)
            self.symbol(k)
            self.item(v)
        self.int(NILVALUE_SXP)

    def strvec(self, strings, attrs=None):
        self.int(STRSXP | (0x200 if attrs else 0))
        self.int(len(strings))
        for s in strings:
            self.charsxp(s)
        if attrs:
            self.attributes(attrs)

    def item(self, obj, extra=None):
        attrs = dict(extra or {})
        if isinstance(obj, RObject):
            attrs.update(obj.attributes)
            obj = obj.value
        if obj is None:
            self.int(NILVALUE_SXP)
            return
        if isinstance(obj, dict):
            attrs = dict(names=list(obj.keys()), **attrs)
            obj = list(obj.values())
        if isinstance(obj, str):
            obj = [obj]
        if isinstance(obj, (list, tuple)):
            if len(obj) and all(isinstance(x, str) or x is None for x in obj):
                self.strvec(list(obj), attrs)
                return
            self.int(VECSXP | (0x200 if attrs else 0))
            self.int(len(obj))
            for x in obj:
                self.item(x)
            if attrs:
                self.attributes(attrs)
            return
        a = np.asarray(obj)
        if a.ndim >= 2:
            attrs = dict(dim=np.asarray(a.shape, dtype=np.int32), **attrs)
        flat = a.reshape(-1, order="F")
        if a.dtype == bool:
            typ, payload = LGLSXP, flat.astype(">i4").tobytes()
        elif np.issubdtype(a.dtype, np.integer):
            typ, payload = INTSXP, flat.astype(">i4").tobytes()
        elif np.issubdtype(a.dtype, np.floating):
            typ, payload = REALSXP, flat.astype(">f8").tobytes()
        else:
            raise TypeError("cannot serialise %r" % (a.dtype,))
        self.int(typ | (0x200 if attrs else 0))
        self.int(flat.size)
        self.out.append(payload)
        if attrs:
            self.attributes(attrs)


def write_rds(path, obj, compress=True, colnames=None, version=2):
    """saveRDS(obj, path, version = 2 | 3, compress = TRUE | "gzip" | "bzip2" | "xz" | FALSE), XDR.  `colnames` adds dimnames(list(NULL, colnames)) to a matrix, the form the
    reference's DIC helpers index their traces by (R/sourceme.R:532)."""
    w = _Writer()
    w.out.append(b"X\n")
    w.int(version)
    w.int(0x00030500)  # written "by" R 3.5.0
    w.int(0x00020300 if version == 2 else 0x00030500)  # minimal R version that reads it
    if version == 3:   # format 3 (R >= 3.5.0 default) adds the native encoding
        w.int(5)
        w.out.append(b"UTF-8")
    elif version != 2:
        raise ValueError("RDS version must be 2 or 3")
    extra = None
    if colnames is not None:
        extra = {"dimnames": [None, list(colnames)]}
    w.item(obj, extra)
    data = b"".join(w.out)
    if compress in (True, "gzip"):
        data = gzip.compress(data)
    elif compress == "bzip2":
        data = bz2.compress(data)
    elif compress == "xz":
        data = lzma.compress(data)
    with open(path, "wb") as f:
        f.write(data)
