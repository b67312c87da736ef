"""GAM reader: decode the vg.Alignment messages this pipeline and the reference write
(framing: concatenated gzip members, each `varint64 count, {varint32 size, message}*`,
src/stream.hpp:24-51; fields: src/vg.proto:52-126).  Used for decoded-message parity --
GAM files are not byte-comparable (record order = completion order, zlib versions)."""
from __future__ import annotations

import struct
import zlib


def _varint(buf, pos):
    r = 0
    shift = 0
    while True:
        c = buf[pos]
        pos += 1
        r |= (c & 0x7F) << shift
        if not c & 0x80:
            return r, pos
        shift += 7


def _fields(buf):
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        f, w = tag >> 3, tag & 7
        if w == 0:
            v, pos = _varint(buf, pos)
        elif w == 1:
            v = struct.unpack_from("<d", buf, pos)[0]
            pos += 8
        elif w == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif w == 5:
            v = struct.unpack_from("<f", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("bad wire type")
        yield f, w, v


def decode_alignment(msg: bytes) -> dict:
    a = dict(sequence="", name="", score=0, query_position=0, identity=0.0, mappings=[])
    for f, w, v in _fields(msg):
        if f == 1:
            a["sequence"] = v.decode()
        elif f == 3:
            a["name"] = v.decode()
        elif f == 6:
            a["score"] = v
        elif f == 7:
            a["query_position"] = v
        elif f == 16:
            a["identity"] = v
        elif f == 2:
            for f2, w2, v2 in _fields(v):
                if f2 != 2:
                    continue
                m = dict(node_id=0, offset=0, is_reverse=False, name="", edits=[], rank=0)
                for f3, w3, v3 in _fields(v2):
                    if f3 == 1:
                        for f4, w4, v4 in _fields(v3):
                            if f4 == 1:
                                m["node_id"] = v4
                            elif f4 == 2:
                                m["offset"] = v4
                            elif f4 == 4:
                                m["is_reverse"] = bool(v4)
                            elif f4 == 5:
                                m["name"] = v4.decode()
                    elif f3 == 2:
                        e = dict(from_length=0, to_length=0, sequence="")
                        for f4, w4, v4 in _fields(v3):
                            if f4 == 1:
                                e["from_length"] = v4
                            elif f4 == 2:
                                e["to_length"] = v4
                            elif f4 == 3:
                                e["sequence"] = v4.decode()
                        m["edits"].append(e)
                    elif f3 == 5:
                        m["rank"] = v3
                a["mappings"].append(m)
    return a


def read_gam(path: str) -> dict:
    """Return {read name: [alignment dict, ...]} (alignments of a read in file order)."""
    with open(path, "rb") as f:
        data = f.read()
    out: dict = {}
    pos = 0
    while pos < len(data):
        d = zlib.decompressobj(31)
        raw = d.decompress(data[pos:])
        pos = len(data) - len(d.unused_data)
        p = 0
        count, p = _varint(raw, p)
        for _ in range(count):
            ln, p = _varint(raw, p)
            a = decode_alignment(raw[p:p + ln])
            p += ln
            out.setdefault(a["name"], []).append(a)
    return out


def message_name(msg: bytes) -> str:
    """`name` (field 3) of a serialized vg.Alignment without decoding the rest (top-level field walk)."""
    pos, n = 0, len(msg)
    while pos < n:
        tag, pos = _varint(msg, pos)
        f, w = tag >> 3, tag & 7
        if w == 0:
            _, pos = _varint(msg, pos)
        elif w == 1:
            pos += 8
        elif w == 5:
            pos += 4
        elif w == 2:
            ln, pos = _varint(msg, pos)
            if f == 3:
                return bytes(msg[pos:pos + ln]).decode()
            pos += ln
        else:
            raise ValueError("bad wire type")
    return ""


def read_gam_messages(data) -> dict:
    """{read name: [serialized message bytes, ...]} of a GAM byte string (messages of a read in file order)."""
    data = bytes(data)
    out: dict = {}
    pos = 0
    while pos < len(data):
        d = zlib.decompressobj(31)
        raw = d.decompress(data[pos:])
        pos = len(data) - len(d.unused_data)
        p = 0
        count, p = _varint(raw, p)
        for _ in range(count):
            ln, p = _varint(raw, p)
            m = raw[p:p + ln]
            p += ln
            out.setdefault(message_name(m), []).append(m)
    return out


def diff_messages(a: dict, b: dict, names=None, limit: int = 5):
    """Compare two {name: [message bytes]} maps over `names` (default: all of both).  Byte-equal messages are equal;
    the rest are decoded and compared field by field.  Returns (reads compared, list of differences)."""
    names = sorted(set(a) | set(b)) if names is None else list(names)
    diffs = []
    for name in names:
        x, y = a.get(name), b.get(name)
        if x == y:
            continue
        if x is None or y is None:
            diffs.append(f"{name}: present only in {'first' if x is not None else 'second'}")
        else:
            dx, dy = {name: [decode_alignment(m) for m in x]}, {name: [decode_alignment(m) for m in y]}
            diffs.extend(diff_gam(dx, dy, limit=1))
        if len(diffs) >= limit:
            break
    return len(names), diffs


def diff_gam(a: dict, b: dict, limit: int = 5):
    """Differences between two decoded GAMs; empty list = identical."""
    diffs = []
    for name in sorted(set(a) | set(b)):
        if name not in a or name not in b:
            diffs.append(f"{name}: present only in {'first' if name in a else 'second'}")
        elif a[name] != b[name]:
            x, y = a[name], b[name]
            if len(x) != len(y):
                diffs.append(f"{name}: {len(x)} vs {len(y)} alignments")
            else:
                for i, (p, q) in enumerate(zip(x, y)):
                    if p != q:
                        keys = [k for k in p if p[k] != q[k]]
                        detail = ""
                        if "mappings" in keys:
                            for mi, (m1, m2) in enumerate(zip(p["mappings"], q["mappings"])):
                                if m1 != m2:
                                    detail = f" first differing mapping {mi}: {m1} vs {m2}"
                                    break
                            else:
                                detail = f" mapping count {len(p['mappings'])} vs {len(q['mappings'])}"
                        diffs.append(f"{name}[{i}]: fields {keys} differ{detail}"[:600])
                        break
        if len(diffs) >= limit:
            break
    return diffs


if __name__ == "__main__":
    import sys
    A, B = read_gam(sys.argv[1]), read_gam(sys.argv[2])
    d = diff_gam(A, B, limit=20)
    print(f"{len(A)} vs {len(B)} reads; {'IDENTICAL' if not d else 'DIFFERENT'}")
    for x in d:
        print("  ", x)
    sys.exit(1 if d else 0)
