"""Minimal BIFF8 (.xls) reader for genedata.xls — the image has no xlrd, and the reference reads the workbook through
pandas.read_excel (bittner/utils.py:10-39).  Parses the OLE2 container (header, DIFAT/FAT, directory), the Workbook
stream's BOUNDSHEET / SST(+CONTINUE) records and the cell records NUMBER, RK, MULRK, LABELSST, LABEL of a sheet."""
import struct
from collections import defaultdict

import numpy as np


def _workbook_stream(data: bytes) -> bytes:
    if data[:8] != bytes.fromhex("D0CF11E0A1B11AE1"):
        raise ValueError("not an OLE2 compound document")
    ssz = 1 << struct.unpack_from("<H", data, 30)[0]
    n_fat = struct.unpack_from("<I", data, 44)[0]
    dir_start = struct.unpack_from("<I", data, 48)[0]
    difat = list(struct.unpack_from("<109I", data, 76))
    nxt, n_difat = struct.unpack_from("<II", data, 68)

    def sector(i):
        return data[512 + i * ssz: 512 + (i + 1) * ssz]

    while n_difat and nxt < 0xFFFFFFFE:
        d = struct.unpack(f"<{ssz // 4}I", sector(nxt))
        difat += d[:-1]
        nxt, n_difat = d[-1], n_difat - 1
    fat = []
    for f in difat[:n_fat]:
        fat += struct.unpack(f"<{ssz // 4}I", sector(f))

    def chain(s):
        out = []
        while s < 0xFFFFFFFE:
            out.append(s)
            s = fat[s]
        return out

    directory = b"".join(sector(s) for s in chain(dir_start))
    for i in range(len(directory) // 128):
        entry = directory[i * 128:(i + 1) * 128]
        name_len = struct.unpack_from("<H", entry, 64)[0]
        name = entry[:max(name_len - 2, 0)].decode("utf-16le")
        start, size = struct.unpack_from("<II", entry, 116)
        if name in ("Workbook", "Book"):
            return b"".join(sector(s) for s in chain(start))[:size]
    raise ValueError("no Workbook stream")


def _rk(v):
    if v & 2:
        x = float((v >> 2) - (1 << 30) if v & 0x80000000 else v >> 2)
    else:
        x = struct.unpack("<d", struct.pack("<II", 0, v & 0xFFFFFFFC))[0]
    return x / 100 if v & 1 else x


class Workbook:
    def __init__(self, path):
        self.wb = _workbook_stream(open(path, "rb").read())
        self.records, pos = [], 0
        while pos + 4 <= len(self.wb):
            rt, rl = struct.unpack_from("<HH", self.wb, pos)
            self.records.append((pos, rt, rl))
            pos += 4 + rl
        self.sheets = {}
        for p, rt, _rl in self.records:
            if rt == 0x85:  # BOUNDSHEET
                off = struct.unpack_from("<I", self.wb, p + 4)[0]
                n, flag = self.wb[p + 10], self.wb[p + 11]
                raw = self.wb[p + 12:p + 12 + n * (2 if flag & 1 else 1)]
                self.sheets[raw.decode("utf-16le" if flag & 1 else "latin1")] = off
        self.sst = self._shared_strings()

    def _shared_strings(self):
        try:
            idx = next(i for i, (_p, rt, _l) in enumerate(self.records) if rt == 0xFC)
        except StopIteration:
            return []
        chunks, j = [], idx
        while j == idx or self.records[j][1] == 0x3C:  # SST then its CONTINUE records
            p, _rt, rl = self.records[j]
            chunks.append(self.wb[p + 4:p + 4 + rl])
            j += 1
        unique = struct.unpack_from("<I", chunks[0], 4)[0]
        ci, off, out = 0, 8, []
        for _ in range(unique):
            if off >= len(chunks[ci]):
                ci, off = ci + 1, 0
            buf = chunks[ci]
            n, flags = struct.unpack_from("<H", buf, off)[0], buf[off + 2]
            off += 3
            rich = ext = 0
            if flags & 8:
                rich = struct.unpack_from("<H", buf, off)[0]
                off += 2
            if flags & 4:
                ext = struct.unpack_from("<I", buf, off)[0]
                off += 4
            wide, parts, left = flags & 1, [], n
            while left > 0:
                buf = chunks[ci]
                width = 2 if wide else 1
                take = min((len(buf) - off) // width, left)
                parts.append(buf[off:off + take * width].decode("utf-16le" if wide else "latin1"))
                off, left = off + take * width, left - take
                if left > 0:  # the string continues in the next CONTINUE record, which restates the width flag
                    ci += 1
                    wide, off = chunks[ci][0] & 1, 1
            out.append("".join(parts))
            skip = rich * 4 + ext
            while skip > 0:
                t = min(len(chunks[ci]) - off, skip)
                off, skip = off + t, skip - t
                if skip > 0:
                    ci, off = ci + 1, 0
        return out

    def cells(self, sheet):
        """{row: {col: value}} of one sheet (numbers as float, strings as str)."""
        out, pos = defaultdict(dict), self.sheets[sheet]
        while True:
            rt, rl = struct.unpack_from("<HH", self.wb, pos)
            b = self.wb[pos + 4:pos + 4 + rl]
            pos += 4 + rl
            if rt == 0x0A:  # EOF of the sheet substream
                return out
            if rt == 0x203:
                r, c = struct.unpack_from("<HH", b)
                out[r][c] = struct.unpack_from("<d", b, 6)[0]
            elif rt == 0x27E:
                r, c = struct.unpack_from("<HH", b)
                out[r][c] = _rk(struct.unpack_from("<I", b, 6)[0])
            elif rt == 0xBD:
                r, c0 = struct.unpack_from("<HH", b)
                for k in range((rl - 6) // 6):
                    out[r][c0 + k] = _rk(struct.unpack_from("<I", b, 4 + 6 * k + 2)[0])
            elif rt == 0xFD:
                r, c = struct.unpack_from("<HH", b)
                out[r][c] = self.sst[struct.unpack_from("<I", b, 6)[0]]
            elif rt == 0x204:
                r, c = struct.unpack_from("<HH", b)
                n = struct.unpack_from("<H", b, 6)[0]
                out[r][c] = b[9:9 + n].decode("latin1")


def read_gene_data(path):
    """(ids int64 [R], names [R], ratios float64 [R][31], weight_ids list) — what extract_gene_data (bittner/utils.py:10-39)
    hands on: sheet "CUTANEOUS MELANOMA" without its last 5 SHEET rows (blank rows count, as they do for
    pandas.read_excel(skipfooter=5)), the 12 "unclustered" + 19 "cluster" ratio columns as T1..T31 (missing cells = NaN),
    and the "Image Clone ID" column of sheet "WEIGHTED GENE LIST"."""
    wb = Workbook(path)
    rows = wb.cells("CUTANEOUS MELANOMA")
    top, sub = rows[0], rows[1]
    id_col = next(c for c, v in sub.items() if v == "Image Clone ID")
    name_col = next(c for c, v in sub.items() if v == "UniGene Cluster Title")
    groups = sorted((c, v) for c, v in top.items() if isinstance(v, str) and v.startswith("Ratio Data"))
    all_cols = sorted(sub)
    cols = []
    for wanted in ("Ratio Data for Group of 12 Unclustered Cutaneous Melanomas", "Ratio Data for Cluster of 19 Cutaneous Melanomas"):
        start = next(c for c, v in groups if v == wanted)
        later = [c for c, _v in groups if c > start]
        stop = min(later) if later else max(all_cols) + 1
        cols += [c for c in all_cols if start <= c < stop]
    data_rows = [r for r in range(2, max(rows) + 1 - 5) if r in rows and id_col in rows[r]]  # skipfooter=5
    ids = np.array([int(rows[r][id_col]) for r in data_rows], dtype=np.int64)
    names = [rows[r].get(name_col, "") for r in data_rows]
    ratios = np.array([[rows[r].get(c, np.nan) for c in cols] for r in data_rows], dtype=np.float64)
    wrows = wb.cells("WEIGHTED GENE LIST")
    wcol = next(c for c, v in wrows[1].items() if v == "Image Clone ID")
    weight_ids = [int(wrows[r][wcol]) for r in sorted(wrows) if r >= 2 and wcol in wrows[r] and not isinstance(wrows[r][wcol], str)]
    return ids, names, ratios, weight_ids
