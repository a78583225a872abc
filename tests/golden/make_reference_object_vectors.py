"""Generates tests/golden/reference_object_vectors.json from the REFERENCE'S OWN OBJECT CODE
(oracle/_ref/libotbref.so, built by oracle/ref/Makefile from /root/reference's leaf sources).
Run in the dev container, where /root/reference exists:

    python -c "import __graft_entry__ as e; e.build()"      # builds oracle/_ref
    python tests/golden/make_reference_object_vectors.py

The vectors travel with the repository, so the oracle stays pinned to values the reference
computed even where neither /root/reference nor oracle/_ref exist
(tests/test_oracle.py::test_oracle_matches_committed_reference_object_vectors)."""
import ctypes as C
import json
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libotbref.so"))
u32, i32, i64, dbl = C.c_uint32, C.c_int32, C.c_int64, C.c_double
for n, res, args in [("ref_hashint4", u32, [i32]), ("ref_hashint8", u32, [i64]), ("ref_hashint4new", u32, [i32]),
                     ("ref_hashint8new", u32, [i64]), ("ref_hashchar", u32, [C.c_int8]), ("ref_hashfloat8", u32, [dbl]),
                     ("ref_murmurhash32", u32, [u32]), ("ref_hash_combine", u32, [u32, u32]), ("ref_hash_any", u32, [C.c_char_p, C.c_int]),
                     ("ref_hash_any_new", u32, [C.c_char_p, C.c_int]),
                     ("ref_evaluate_hashkey1", u32, [C.c_int, i64, C.c_int]), ("ref_evaluate_hashkey2", u32, [i64, i32]),
                     ("ref_float8_accum", None, [C.c_void_p, dbl]), ("ref_float8pl", dbl, [dbl, dbl])]:
    f = getattr(R, n); f.restype = res; f.argtypes = args

rng = np.random.default_rng(20240922)
i4 = [0, 1, -1, 17, 42, 550273, 207112489, 2**31 - 1, -2**31] + rng.integers(-2**31, 2**31, 55).tolist()
i8 = [0, 1, -1, 2**32 + 1, -2**32, 6000000000, 2**63 - 1, -2**63] + rng.integers(-2**63, 2**63 - 1, 56).tolist()
f8 = [0.0, -0.0, 1.0, -1.5, 901.0, 1e300, float("inf")] + rng.normal(0, 1e6, 25).tolist()
blobs = [bytes(rng.integers(0, 256, n).astype(np.uint8)) for n in (0, 1, 3, 4, 7, 8, 11, 12, 13, 24, 31, 64)]
vals = rng.normal(1000, 300, 200).tolist()
state = (dbl * 3)(0.0, 0.0, 0.0)
s = None
for x in vals:
    R.ref_float8_accum(state, x)
    s = x if s is None else R.ref_float8pl(s, x)
out = {
    "source": "oracle/_ref/libotbref.so = hashfunc.c, pg_crc32c_sb8.c, float.c, locator.c of the reference, compiled in place",
    "hashint4": [[v, R.ref_hashint4(v)] for v in i4],
    "hashint4new": [[v, R.ref_hashint4new(v)] for v in i4],
    "hashint8": [[v, R.ref_hashint8(v)] for v in i8],
    "hashint8new": [[v, R.ref_hashint8new(v)] for v in i8],
    "murmurhash32": [[v & 0xFFFFFFFF, R.ref_murmurhash32(v & 0xFFFFFFFF)] for v in i4],
    "hashchar": [[v, R.ref_hashchar(v)] for v in (-128, -1, 0, 65, 78, 82, 127)],
    "hashfloat8_bits": [[struct.unpack("<q", struct.pack("<d", v))[0], R.ref_hashfloat8(v)] for v in f8],
    "hash_combine": [[a, b, R.ref_hash_combine(a, b)] for a, b in rng.integers(0, 2**32, (16, 2)).tolist()],
    "hash_any": [[b.hex(), R.ref_hash_any(b, len(b)), R.ref_hash_any_new(b, len(b))] for b in blobs],
    "evaluate_hashkey_int4": [[v, R.ref_evaluate_hashkey1(0, v, 0)] for v in i4],
    "evaluate_hashkey_int8": [[v, R.ref_evaluate_hashkey1(1, v, 0)] for v in i8],
    "evaluate_hashkey_int8_int4": [[a, b, R.ref_evaluate_hashkey2(a, b)] for a, b in zip(i8[:24], i4[:24])],
    "float8_accum": {"inputs_bits": [struct.unpack("<q", struct.pack("<d", v))[0] for v in vals],
                     "state_bits": [struct.unpack("<q", struct.pack("<d", v))[0] for v in list(state)],
                     "float8pl_sum_bits": struct.unpack("<q", struct.pack("<d", s))[0]},
}
path = os.path.join(ROOT, "tests", "golden", "reference_object_vectors.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, os.path.getsize(path), "bytes")
