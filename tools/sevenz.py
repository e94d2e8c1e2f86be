#!/usr/bin/env python
"""Minimal 7z reader for the reference's bundled fixtures (datasets/texts.7z, datasets/patterns.7z).

No 7z tool exists in this image; both archives are 7z v0.3 containers with an LZMA1-encoded header
and ONE solid LZMA1 folder holding the files as sub-streams. stdlib `lzma` decodes
the raw LZMA1 streams; this module parses just enough of the container: signature header, encoded
header, PackInfo / UnpackInfo / SubStreamsInfo, FilesInfo names and empty-stream flags. Every
extracted file is CRC32-checked against the archive's own digest.

    python tools/sevenz.py list  /root/reference/datasets/patterns.7z
    python tools/sevenz.py get   /root/reference/datasets/texts.7z world_leaders out.txt
"""
import lzma
import struct
import sys
import zlib


def _num(b, p):
    first = b[p]; p += 1
    mask, value = 0x80, 0
    for i in range(8):
        if first & mask == 0:
            return value | ((first & (mask - 1)) << (8 * i)), p
        value |= b[p] << (8 * i); p += 1
        mask >>= 1
    return value, p


def _bits(b, p, n):
    out = []
    for i in range(n):
        out.append(bool(b[p + i // 8] & (0x80 >> (i % 8))))
    return out, p + (n + 7) // 8


def _parse_streams_info(b, p):
    """Returns (info dict, new position). Handles one pack stream per folder, one coder per folder."""
    info = {"pack_pos": 0, "pack_sizes": [], "folders": [], "unpack_sizes": [], "sub_counts": None,
            "sub_sizes": None, "sub_crcs": None}
    while True:
        t = b[p]; p += 1
        if t == 0x00:
            return info, p
        if t == 0x06:  # PackInfo
            info["pack_pos"], p = _num(b, p)
            n, p = _num(b, p)
            while True:
                t2 = b[p]; p += 1
                if t2 == 0x00:
                    break
                if t2 == 0x09:
                    for _ in range(n):
                        s, p = _num(b, p); info["pack_sizes"].append(s)
                elif t2 == 0x0A:
                    alld = b[p]; p += 1
                    defined = [True] * n
                    if not alld:
                        defined, p = _bits(b, p, n)
                    p += 4 * sum(defined)
                else:
                    raise ValueError("PackInfo property %#x" % t2)
        elif t == 0x07:  # UnpackInfo
            assert b[p] == 0x0B; p += 1
            nf, p = _num(b, p)
            assert b[p] == 0; p += 1  # not external
            for _ in range(nf):
                nc, p = _num(b, p)
                assert nc == 1, "only single-coder folders are supported"
                flag = b[p]; p += 1
                idsz = flag & 0x0F
                cid = bytes(b[p:p + idsz]); p += idsz
                assert not (flag & 0x10), "complex coders are not supported"
                props = b""
                if flag & 0x20:
                    ps, p = _num(b, p)
                    props = bytes(b[p:p + ps]); p += ps
                info["folders"].append({"id": cid, "props": props})
            assert b[p] == 0x0C; p += 1
            for _ in range(nf):
                s, p = _num(b, p); info["unpack_sizes"].append(s)
            t2 = b[p]; p += 1
            if t2 == 0x0A:
                alld = b[p]; p += 1
                defined = [True] * nf
                if not alld:
                    defined, p = _bits(b, p, nf)
                info["folder_crcs"] = []
                for d in defined:
                    if d:
                        info["folder_crcs"].append(struct.unpack_from("<I", b, p)[0]); p += 4
                    else:
                        info["folder_crcs"].append(None)
                t2 = b[p]; p += 1
            assert t2 == 0x00
        elif t == 0x08:  # SubStreamsInfo
            nf = len(info["folders"])
            counts = [1] * nf
            sizes = None
            crcs = None
            while True:
                t2 = b[p]; p += 1
                if t2 == 0x00:
                    break
                if t2 == 0x0D:
                    counts = []
                    for _ in range(nf):
                        c, p = _num(b, p); counts.append(c)
                elif t2 == 0x09:
                    sizes = []
                    for f in range(nf):
                        acc, cur = 0, []
                        for _ in range(counts[f] - 1):
                            s, p = _num(b, p); cur.append(s); acc += s
                        cur.append(info["unpack_sizes"][f] - acc)
                        sizes.append(cur)
                elif t2 == 0x0A:
                    total = sum(counts)
                    alld = b[p]; p += 1
                    defined = [True] * total
                    if not alld:
                        defined, p = _bits(b, p, total)
                    crcs = []
                    for d in defined:
                        if d:
                            crcs.append(struct.unpack_from("<I", b, p)[0]); p += 4
                        else:
                            crcs.append(None)
                else:
                    raise ValueError("SubStreamsInfo property %#x" % t2)
            if sizes is None:
                sizes = [[info["unpack_sizes"][f]] if counts[f] == 1 else None for f in range(nf)]
            info["sub_counts"], info["sub_sizes"], info["sub_crcs"] = counts, sizes, crcs
        else:
            raise ValueError("StreamsInfo property %#x" % t)


def _lzma1_decoder(props):
    d = props[0]
    lc, d = d % 9, d // 9
    lp, pb = d % 5, d // 5
    dict_size = struct.unpack("<I", props[1:5])[0]
    return lzma.LZMADecompressor(format=lzma.FORMAT_RAW,
                                 filters=[{"id": lzma.FILTER_LZMA1, "lc": lc, "lp": lp, "pb": pb, "dict_size": max(dict_size, 4096)}])


class Archive:
    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            sig = f.read(32)
            assert sig[:6] == b"7z\xbc\xaf\x27\x1c", "not a 7z file"
            off, size, crc = struct.unpack("<QQI", sig[12:32])
            f.seek(32 + off)
            hdr = f.read(size)
            assert zlib.crc32(hdr) == crc, "next-header CRC mismatch"
            if hdr[0] == 0x17:  # EncodedHeader: the real header is itself a packed stream
                info, _ = _parse_streams_info(hdr, 1)
                assert info["folders"][0]["id"] == b"\x03\x01\x01", "header coder is not LZMA1"
                f.seek(32 + info["pack_pos"])
                packed = f.read(info["pack_sizes"][0])
                hdr = _lzma1_decoder(info["folders"][0]["props"]).decompress(packed, max_length=info["unpack_sizes"][0])
                assert len(hdr) == info["unpack_sizes"][0]
        assert hdr[0] == 0x01, "Header expected"
        p = 1
        self.streams = None
        self.names, self.empty = [], []
        while True:
            t = hdr[p]; p += 1
            if t == 0x00:
                break
            if t == 0x04:
                self.streams, p = _parse_streams_info(hdr, p)
            elif t == 0x05:
                nfiles, p = _num(hdr, p)
                self.empty = [False] * nfiles
                while True:
                    pt = hdr[p]; p += 1
                    if pt == 0x00:
                        break
                    psz, p = _num(hdr, p)
                    body = hdr[p:p + psz]; p += psz
                    if pt == 0x0E:
                        self.empty, _ = _bits(body, 0, nfiles)
                    elif pt == 0x11:
                        assert body[0] == 0
                        self.names = [s for s in bytes(body[1:]).decode("utf-16-le").split("\0") if s != "" or False][:nfiles]
            else:
                raise ValueError("header property %#x" % t)
        assert self.streams is not None and len(self.streams["folders"]) == 1, "one solid folder expected"
        assert self.streams["folders"][0]["id"] == b"\x03\x01\x01", "payload coder is not LZMA1"
        files = [n for n, e in zip(self.names, self.empty) if not e]
        sizes = self.streams["sub_sizes"][0]
        crcs = self.streams["sub_crcs"] or [None] * len(sizes)
        assert len(files) == len(sizes)
        self.files = []
        off = 0
        for n, s, c in zip(files, sizes, crcs):
            self.files.append({"name": n, "size": s, "crc": c, "offset": off})
            off += s

    def extract(self, name):
        """Bytes of the member whose basename is `name` (streams the solid folder up to it)."""
        ent = next(f for f in self.files if f["name"].split("/")[-1] == name)
        dec = _lzma1_decoder(self.streams["folders"][0]["props"])
        out = bytearray()
        skip, want = ent["offset"], ent["size"]
        with open(self.path, "rb") as f:
            f.seek(32 + self.streams["pack_pos"])
            remaining_in = self.streams["pack_sizes"][0]
            while len(out) < want and not dec.eof:
                chunk = b""
                if dec.needs_input and remaining_in > 0:
                    chunk = f.read(min(1 << 20, remaining_in))
                    remaining_in -= len(chunk)
                piece = dec.decompress(chunk, max_length=1 << 24)
                if not piece and not chunk and remaining_in == 0:
                    break  # raw LZMA1 in 7z has no end marker: the stream simply runs out
                if skip:
                    d = min(skip, len(piece))
                    piece = piece[d:]; skip -= d
                out += piece[: want - len(out)]
        assert len(out) == want, "short read"
        if ent["crc"] is not None:
            assert zlib.crc32(out) == ent["crc"], "CRC mismatch for " + name
        return bytes(out)


def extract_one(archive, name):
    return Archive(archive).extract(name)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "list":
        for f in Archive(sys.argv[2]).files:
            print("%12d  %08x  %s" % (f["size"], f["crc"] or 0, f["name"]))
    elif len(sys.argv) >= 5 and sys.argv[1] == "get":
        open(sys.argv[4], "wb").write(extract_one(sys.argv[2], sys.argv[3]))
    else:
        print(__doc__)
