"""Extracts the columns BASELINE config 1 needs from the reference's shipped data set
(/root/reference/data/pbc2.rda, pbc2.id.rda: gzip'd R serialisation, format RDX2 / XDR) into
tests/golden/pbc2_columns.npz.  Run in the build container, where the reference tree is mounted:

    python tests/golden/make_pbc2.py

config 1 = mvglmer(log(serBilir) ~ year + (year | id)) + coxph(Surv(years, status2) ~ drug + age)
(man/mvJointModelBayes.Rd:204-227).  Only numeric columns are kept: id, year, serBilir (long format,
1945 rows) and years, status2, drug (0 = placebo, 1 = D-penicil), age (one row per subject, 312 rows).
"""
import gzip
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("JMB_REFERENCE", "/root/reference")


class RDX2:
    """Minimal reader of R's XDR serialisation (version 2): enough for data frames."""

    def __init__(self, raw: bytes):
        assert raw[:5] == b"RDX2\n" and raw[5:7] == b"X\n", "not an RDX2/XDR file"
        self.b, self.p, self.refs = raw, 7, []
        self.p += 12                                   # version, writer version, min reader version

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def item(self):
        flags = self.i32()
        t, has_attr, has_tag = flags & 0xFF, bool(flags & 0x200), bool(flags & 0x400)
        if t == 254:                                   # NILVALUE
            return None
        if t == 255:                                   # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:                                     # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:                                     # LISTSXP (pairlist): dict of tag -> value
            out = {}
            while True:
                attr = self.item() if has_attr else None   # noqa: F841
                tag = self.item() if has_tag else None
                out[tag if tag is not None else len(out)] = self.item()
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == 254:
                    return out
                assert t2 == 2, t2
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 9:                                     # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("latin-1")
            self.p += n
            return s
        if t in (10, 13):                              # LGLSXP / INTSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int64)
            self.p += 4 * n
            res = v
        elif t == 14:                                  # REALSXP
            n = self.i32()
            res = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
        elif t == 16:                                  # STRSXP
            n = self.i32()
            res = [self.item() for _ in range(n)]
        elif t == 19:                                  # VECSXP
            n = self.i32()
            res = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {t}")
        attrs = self.item() if has_attr else None
        if attrs:
            return {"value": res, "attr": attrs}
        return res


def read_rda(path):
    r = RDX2(gzip.open(path, "rb").read())
    top = r.item()                                     # pairlist: object name -> data frame
    name, df = next(iter(top.items()))
    cols = dict(zip(df["attr"]["names"], df["value"]))
    return name, cols


def col(c):
    """numeric view of a column: factors -> integer codes (1-based) with levels."""
    if isinstance(c, dict):
        return np.asarray(c["value"]), c["attr"].get("levels")
    return np.asarray(c), None


if __name__ == "__main__":
    _, long_ = read_rda(os.path.join(REF, "data", "pbc2.rda"))
    _, wide = read_rda(os.path.join(REF, "data", "pbc2.id.rda"))
    ids, id_levels = col(long_["id"])
    year, _ = col(long_["year"])
    bili, _ = col(long_["serBilir"])
    years, _ = col(wide["years"])
    status2, _ = col(wide["status2"])
    drug, drug_levels = col(wide["drug"])
    age, _ = col(wide["age"])
    wid, _ = col(wide["id"])
    assert len(ids) == 1945 and len(years) == 312 and drug_levels == ["placebo", "D-penicil"]
    np.savez_compressed(os.path.join(HERE, "pbc2_columns.npz"),
                        id=ids.astype(np.int32) - 1, year=year, serBilir=bili,
                        years=years, status2=status2.astype(np.float64), drug=(drug == 2).astype(np.float64), age=age,
                        wide_id=wid.astype(np.int32) - 1)
    print("wrote pbc2_columns.npz:", len(ids), "rows,", len(years), "subjects; events:", int(status2.sum()))
