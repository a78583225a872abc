"""Generates tests/golden/fnpage_vectors.json from the REFERENCE'S OWN OBJECT CODE: forward-node pages (the
redistribute wire format) produced by heaptuple.o + fnbufpage.o through oracle/ref/ref_glue.c::ref_fnpage_pack,
which follows FragmentSendAttrs / FragmentGetPage / FragmentSendNullTuple (executor/execFragment.c:2067-2136,
:1857-1876, :1963-1975).  Run in the dev container, where /root/reference exists:

    python -c "import __graft_entry__ as e; e.build()"      # builds oracle/_ref
    python tests/golden/make_fnpage_vectors.py

Only the used part of every page ([0, lower)) is stored; the rest of a page is zero in the fixture's convention."""
import ctypes as C
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libotbref.so"))
R.ref_fnpage_pack.restype = C.c_int64
R.ref_fnpage_pack.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
# gpuexec.h type codes -> (attlen, attalign); 6 = bpchar(1), a short varlena on the wire
ATT = {1: (4, 4), 2: (8, 8), 3: (8, 8), 4: (4, 4), 5: (1, 1), 6: (-1, 4)}
rng = np.random.default_rng(20240923)


def case(name, types, n, null_frac, page_id, end_marker=True):
    vals = np.zeros((n, len(types)), np.int64)
    for i, t in enumerate(types):
        if t == 2:
            vals[:, i] = rng.integers(-2**62, 2**62, n)
        elif t == 3:
            vals[:, i] = rng.normal(0, 1e6, n).view(np.int64)
        elif t in (1, 4):
            vals[:, i] = rng.integers(-2**31, 2**31 - 1, n)
        else:
            vals[:, i] = rng.integers(32, 127, n)
    isnull = (rng.random((n, len(types))) < null_frac).astype(np.uint8)
    isnull[: max(1, n // 8)] = 0                                   # some tuples without a bitmap even in the nullable cases
    attlen = (C.c_int16 * len(types))(*[ATT[t][0] for t in types])
    attalign = (C.c_int8 * len(types))(*[ATT[t][1] for t in types])
    cap = n // 20 + 4
    pages = np.zeros((cap, 8192), np.uint8)
    k = R.ref_fnpage_pack(len(types), attlen, attalign, vals.ctypes.data, isnull.ctypes.data, n, page_id["qid_ts"], page_id["qid_seq"],
                          page_id["fid"], page_id["nodeid"], page_id["workerid"], page_id["virtualid"], int(end_marker), pages.ctypes.data, cap)
    assert k > 0
    used = []
    for p in range(k):
        lower = int(pages[p, 0:4].copy().view(np.uint32)[0])
        assert not pages[p, lower:].any()
        used.append(pages[p, :lower].tobytes().hex())
    return {"name": name, "types": types, "values": vals.tolist(), "isnull": isnull.tolist(), "page_id": page_id,
            "end_marker": end_marker, "pages_used_hex": used}


out = {"source": "oracle/_ref/libotbref.so (heaptuple.c, fnbufpage.c of the reference compiled in place) via ref_fnpage_pack",
       "cases": [
           case("q3 orders row, no NULLs", [2, 2, 4, 1], 450, 0.0, dict(qid_ts=0x0123456789ABCDEF, qid_seq=77, fid=3, nodeid=1, workerid=0, virtualid=0)),
           case("mixed widths with NULLs", [2, 1, 3, 5, 4, 5, 2], 260, 0.2, dict(qid_ts=-5, qid_seq=2**40, fid=65535, nodeid=7, workerid=2, virtualid=1)),
           case("bpchar(1) and ten attributes", [1, 6, 6, 3, 2, 5, 4, 1, 6, 2], 200, 0.15, dict(qid_ts=1, qid_seq=1, fid=1, nodeid=0, workerid=0, virtualid=0)),
           case("exactly full last page, no end marker", [2, 2, 2], 204, 0.0, dict(qid_ts=9, qid_seq=9, fid=9, nodeid=9, workerid=9, virtualid=9), end_marker=False),
           case("end marker alone on a fresh page", [2, 2, 2], 204, 0.0, dict(qid_ts=9, qid_seq=9, fid=9, nodeid=9, workerid=9, virtualid=9)),
       ]}
path = os.path.join(ROOT, "tests", "golden", "fnpage_vectors.json")
json.dump(out, open(path, "w"))
print("wrote", path, os.path.getsize(path), "bytes;", [len(c["pages_used_hex"]) for c in out["cases"]], "pages")
