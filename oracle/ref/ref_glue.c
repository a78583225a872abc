/*
 * ref_glue.c — thin plain-C doors into the REFERENCE's own compiled functions.
 * TEST INFRASTRUCTURE (oracle/_ref/libotbref.so).  Compiled against the
 * reference's real headers; everything computed here is computed by reference
 * object code (hashfunc.o, pg_crc32c_sb8.o, bloomfilter.o, heaptuple.o,
 * bufpage.o, float.o, int8.o, locator.o).  What this file adds is only what a
 * stand-alone process lacks: palloc (-> malloc), elog (-> abort), the fmgr
 * call trampoline, and a hand-made TupleDesc / float8[3] array.  Same trick as
 * the reference's own unit tests (src/backend/unittest/backend/stub).
 */
#include "postgres.h"

#include "access/hash.h"
#include "access/htup_details.h"
#include "access/tupdesc.h"
#include "catalog/pg_type.h"
#include "fmgr.h"
#include "nodes/execnodes.h"
#include "pgxc/locator.h"
#include "storage/bufpage.h"
#include "utils/array.h"
#include "utils/bloomfilter.h"
#include "utils/builtins.h"
#include "utils/hashutils.h"
#include "forward/fnbufpage.h"
#include "pgxc/squeue.h"          /* MAX_UINT32, the end-of-stream length word */

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- what a backend would provide -------------------------------------- */
/* OpenTenBase's palloc family carries file/line (utils/palloc.h:80-108) */
void *palloc_internal(Size size, const char *file, int line) { return malloc(size ? size : 1); }
void *palloc0_internal(Size size, const char *file, int line) { return calloc(1, size ? size : 1); }
void *palloc_extended_internal(Size size, int flags, const char *file, int line) { return calloc(1, size ? size : 1); }
void *repallocInternal(void *p, Size size, const char *file, int line) { return realloc(p, size); }
void pfree(void *p) { free(p); }
void *MemoryContextAllocInternal(MemoryContext c, Size size, const char *file, int line) { return malloc(size ? size : 1); }
void *MemoryContextAllocZeroInternal(MemoryContext c, Size size, const char *file, int line) { return calloc(1, size ? size : 1); }
void *MemoryContextAllocZeroAlignedIternal(MemoryContext c, Size size, const char *file, int line) { return calloc(1, size ? size : 1); }
__thread MemoryContext CurrentMemoryContext = NULL;
volatile bool InterruptPending = false;
int  ref_elog_count = 0;
bool errstart(int elevel, const char *filename, int lineno, const char *funcname, const char *domain)
{ if (elevel >= ERROR) { ref_elog_count++; fprintf(stderr, "reference ereport(ERROR) at %s:%d %s\n", filename, lineno, funcname); abort(); } return false; }
void errfinish(int dummy,...) { }
static const char *elog_file; static int elog_line;
void elog_start(const char *filename, int lineno, const char *funcname) { elog_file = filename; elog_line = lineno; }
void elog_finish(int elevel, const char *fmt,...) { if (elevel >= ERROR) { fprintf(stderr, "reference elog(ERROR) at %s:%d: %s\n", elog_file, elog_line, fmt); abort(); } }
int errmsg(const char *fmt,...) { return 0; }
int errmsg_internal(const char *fmt,...) { return 0; }
int errcode(int sqlerrcode) { return 0; }
int errdetail(const char *fmt,...) { return 0; }
int errhint(const char *fmt,...) { return 0; }
void ExceptionalCondition(const char *a, const char *b, const char *c, int d) { abort(); }
struct varlena *pg_detoast_datum(struct varlena *datum) { return datum; }
struct varlena *pg_detoast_datum_packed(struct varlena *datum) { return datum; }

/* DirectFunctionCall1Coll, utils/fmgr/fmgr.c:796 */
Datum DirectFunctionCall1Coll(PGFunction func, Oid collation, Datum arg1)
{
    FunctionCallInfoData fcinfo;
    InitFunctionCallInfoData(fcinfo, NULL, 1, collation, NULL, NULL);
    fcinfo.arg[0] = arg1; fcinfo.argnull[0] = false;
    return (*func) (&fcinfo);
}
static Datum call2(PGFunction func, Datum a, Datum b, fmNodePtr ctx)
{
    FunctionCallInfoData fcinfo;
    InitFunctionCallInfoData(fcinfo, NULL, 2, InvalidOid, ctx, NULL);
    fcinfo.arg[0] = a; fcinfo.arg[1] = b; fcinfo.argnull[0] = fcinfo.argnull[1] = false;
    return (*func) (&fcinfo);
}
/* float8_accum/_combine modify their state in place only inside an aggregate */
int AggCheckCallContext(FunctionCallInfo fcinfo, MemoryContext *aggcontext) { return AGG_CONTEXT_AGGREGATE; }

/* ---- hashes -------------------------------------------------------------- */
uint32 ref_hash_any(const unsigned char *k, int len) { return DatumGetUInt32(hash_any(k, len)); }
uint32 ref_hash_uint32(uint32 k) { return DatumGetUInt32(hash_uint32(k)); }
uint32 ref_hashint4(int32 v) { return DatumGetUInt32(DirectFunctionCall1(hashint4, Int32GetDatum(v))); }
uint32 ref_hashint8(int64 v) { return DatumGetUInt32(DirectFunctionCall1(hashint8, Int64GetDatum(v))); }
uint32 ref_hashchar(int8 v) { return DatumGetUInt32(DirectFunctionCall1(hashchar, CharGetDatum(v))); }
uint32 ref_hashfloat8(double v) { return DatumGetUInt32(DirectFunctionCall1(hashfloat8, Float8GetDatum(v))); }
uint32 ref_hash_any_new(const unsigned char *k, int len) { return DatumGetUInt32(hash_any_new(k, len)); }
uint32 ref_hashint4new(int32 v) { return DatumGetUInt32(DirectFunctionCall1(hashint4new, Int32GetDatum(v))); }
uint32 ref_hashint8new(int64 v) { return DatumGetUInt32(DirectFunctionCall1(hashint8new, Int64GetDatum(v))); }
uint32 ref_hashcharnew(int8 v) { return DatumGetUInt32(DirectFunctionCall1(hashcharnew, CharGetDatum(v))); }
uint32 ref_hashfloat8new(double v) { return DatumGetUInt32(DirectFunctionCall1(hashfloat8new, Float8GetDatum(v))); }
uint32 ref_murmurhash32(uint32 h) { return murmurhash32(h); }
uint32 ref_hash_combine(uint32 a, uint32 b) { return hash_combine(a, b); }

/* EvaluateHashkey (pgxc/locator/locator.c:1611) for one int4 / int8 / date column */
uint32 ref_evaluate_hashkey1(int is_int8, int64 v, int isnull)
{
    Oid type = is_int8 ? INT8OID : INT4OID;
    bool n = isnull != 0;
    Datum d = is_int8 ? Int64GetDatum(v) : Int32GetDatum((int32) v);
    return EvaluateHashkey(&type, &n, &d, 1);
}
uint32 ref_evaluate_hashkey2(int64 a_int8, int32 b_int4)
{
    Oid types[2] = { INT8OID, INT4OID };
    bool n[2] = { false, false };
    Datum d[2] = { Int64GetDatum(a_int8), Int32GetDatum(b_int4) };
    return EvaluateHashkey(types, n, d, 2);
}

/* ---- bloom filter --------------------------------------------------------- */
void *ref_bloom_init(double nrows, double fpp) { return BlockBloomFilterInit(nrows, fpp); }
void ref_bloom_insert(void *f, uint32 h) { BlockBloomFilterInsert((BlockBloomFilter) f, h); }
int ref_bloom_find(void *f, uint32 h) { return BlockBloomFilterFind((BlockBloomFilter) f, h, NULL, NULL); }
int ref_bloom_log_num_buckets(void *f) { return f ? ((BlockBloomFilter) f)->logNumBuckets : -1; }
const uint32 *ref_bloom_words(void *f) { return (const uint32 *) ((BlockBloomFilter) f)->directory; }

/* ---- float8 transition / combine / final functions ------------------------ */
static ArrayType *make_f8_array3(const double *v)
{
    Size sz = ARR_OVERHEAD_NONULLS(1) + 3 * sizeof(float8);
    ArrayType *a = (ArrayType *) calloc(1, sz);
    SET_VARSIZE(a, sz);
    a->ndim = 1; a->dataoffset = 0; a->elemtype = FLOAT8OID;
    ARR_DIMS(a)[0] = 3; ARR_LBOUND(a)[0] = 1;
    memcpy(ARR_DATA_PTR(a), v, 3 * sizeof(float8));
    return a;
}
double ref_float8pl(double a, double b) { return DatumGetFloat8(call2(float8pl, Float8GetDatum(a), Float8GetDatum(b), NULL)); }
double ref_float8mul(double a, double b) { return DatumGetFloat8(call2(float8mul, Float8GetDatum(a), Float8GetDatum(b), NULL)); }
double ref_float8mi(double a, double b) { return DatumGetFloat8(call2(float8mi, Float8GetDatum(a), Float8GetDatum(b), NULL)); }
/* state[3] = {N, Sx, Sxx} updated in place by the reference's float8_accum */
void ref_float8_accum(double *state, double newval)
{
    ArrayType *a = make_f8_array3(state);
    call2(float8_accum, PointerGetDatum(a), Float8GetDatum(newval), NULL);
    memcpy(state, ARR_DATA_PTR(a), 3 * sizeof(float8));
    free(a);
}
void ref_float8_combine(double *state1, const double *state2)
{
    ArrayType *a = make_f8_array3(state1), *b = make_f8_array3(state2);
    call2(float8_combine, PointerGetDatum(a), PointerGetDatum(b), NULL);
    memcpy(state1, ARR_DATA_PTR(a), 3 * sizeof(float8));
    free(a); free(b);
}
/* returns 0 and *out, or 1 for SQL NULL */
int ref_float8_avg(const double *state, double *out)
{
    ArrayType *a = make_f8_array3(state);
    FunctionCallInfoData fcinfo;
    InitFunctionCallInfoData(fcinfo, NULL, 1, InvalidOid, NULL, NULL);
    fcinfo.arg[0] = PointerGetDatum(a); fcinfo.argnull[0] = false;
    Datum d = float8_avg(&fcinfo);
    free(a);
    if (fcinfo.isnull) return 1;
    *out = DatumGetFloat8(d);
    return 0;
}
int64 ref_int8inc(int64 v)
{
    FunctionCallInfoData fcinfo;
    InitFunctionCallInfoData(fcinfo, NULL, 1, InvalidOid, NULL, NULL);
    fcinfo.arg[0] = Int64GetDatum(v); fcinfo.argnull[0] = false;
    return DatumGetInt64(int8inc(&fcinfo));
}

/* ---- heap tuples and pages ------------------------------------------------- */
static TupleDesc make_desc(int natts, const int16 *attlen, const int8 *attalign)
{
    TupleDesc d = (TupleDesc) calloc(1, offsetof(struct tupleDesc, attrs) + natts * sizeof(FormData_pg_attribute));
    d->natts = natts; d->tdtypeid = RECORDOID; d->tdtypmod = -1; d->tdrefcount = -1;
    for (int i = 0; i < natts; i++) {
        Form_pg_attribute a = TupleDescAttr(d, i);
        a->attnum = i + 1; a->attlen = attlen[i]; a->attcacheoff = -1; a->atttypmod = -1;
        a->attbyval = attlen[i] > 0;
        a->attalign = attalign[i] == 8 ? 'd' : attalign[i] == 4 ? 'i' : attalign[i] == 2 ? 's' : 'c';
        a->attstorage = attlen[i] > 0 ? 'p' : 'x';
        a->atttypid = attlen[i] == 8 ? INT8OID : attlen[i] == 4 ? INT4OID : attlen[i] == 1 ? CHAROID : BPCHAROID;
    }
    return d;
}
/* values[i]: by-value datum, or for attlen -1 the single character of a bpchar(1).
 * Forms the tuple with the reference's heap_form_tuple and copies its bytes
 * (t_len of them) to out; returns t_len. */
int ref_heap_form_tuple(int natts, const int16 *attlen, const int8 *attalign,
                        const int64 *values, const uint8 *isnull, uint8 *out, int outcap)
{
    TupleDesc d = make_desc(natts, attlen, attalign);
    Datum *v = (Datum *) calloc(natts, sizeof(Datum));
    bool *n = (bool *) calloc(natts, sizeof(bool));
    for (int i = 0; i < natts; i++) {
        n[i] = isnull[i] != 0;
        if (attlen[i] == -1) {
            /* a 4-byte-header varlena holding one character, as bpcharin would make it */
            struct varlena *t = (struct varlena *) calloc(1, VARHDRSZ + 1);
            SET_VARSIZE(t, VARHDRSZ + 1);
            VARDATA(t)[0] = (char) values[i];
            v[i] = PointerGetDatum(t);
        } else v[i] = (Datum) values[i];
    }
    HeapTuple tup = heap_form_tuple(d, v, n);
    int len = (int) tup->t_len;
    if (len <= outcap) memcpy(out, tup->t_data, len);
    free(v); free(n); free(d);
    return len;
}
/* deform `len` tuple bytes with the reference's heap_deform_tuple */
void ref_heap_deform_tuple(int natts, const int16 *attlen, const int8 *attalign,
                           const uint8 *tuple, int len, int64 *values_out, uint8 *isnull_out)
{
    TupleDesc d = make_desc(natts, attlen, attalign);
    HeapTupleData htup;
    Datum *v = (Datum *) calloc(natts, sizeof(Datum));
    bool *n = (bool *) calloc(natts, sizeof(bool));
    void *copy = malloc(len + 8);
    memcpy(copy, tuple, len);
    memset(&htup, 0, sizeof(htup));
    htup.t_len = len; htup.t_data = (HeapTupleHeader) copy;
    heap_deform_tuple(&htup, d, v, n);
    for (int i = 0; i < natts; i++) {
        isnull_out[i] = n[i];
        if (n[i]) { values_out[i] = 0; continue; }
        if (attlen[i] == -1) {
            struct varlena *t = (struct varlena *) DatumGetPointer(v[i]);
            values_out[i] = (int64) (int8) VARDATA_ANY(t)[0];
        } else if (attlen[i] == 1) values_out[i] = (int64) DatumGetChar(v[i]);
        else if (attlen[i] == 2) values_out[i] = (int64) DatumGetInt16(v[i]);
        else if (attlen[i] == 4) values_out[i] = (int64) DatumGetInt32(v[i]);
        else values_out[i] = DatumGetInt64(v[i]);
    }
    free(v); free(n); free(d); free(copy);
}
/* PageInit + PageAddItem of `ntuples` items (lengths lens[], bytes concatenated in items) */
int ref_page_build(uint8 *page, const uint8 *items, const int *lens, int ntuples)
{
    PageInit((Page) page, BLCKSZ, 0, false);
    int off = 0, added = 0;
    for (int i = 0; i < ntuples; i++) {
        if (PageGetFreeSpace((Page) page) < MAXALIGN(lens[i])) break;
        OffsetNumber o = PageAddItem((Page) page, (Item) (items + off), lens[i], InvalidOffsetNumber, false, true);
        if (o == InvalidOffsetNumber) break;
        off += lens[i]; added++;
    }
    return added;
}
int ref_sizeof_heap_header(void) { return (int) SizeofHeapTupleHeader; }
int ref_sizeof_page_header(void) { return (int) SizeOfPageHeaderData; }
int ref_sizeof_minimal_header(void) { return (int) SizeofMinimalTupleHeader; }
int ref_minimal_tuple_offset(void) { return (int) MINIMAL_TUPLE_OFFSET; }
int ref_offsetof_hoff(void) { return (int) offsetof(HeapTupleHeaderData, t_hoff); }
int ref_offsetof_infomask(void) { return (int) offsetof(HeapTupleHeaderData, t_infomask); }
void ref_page_offsets(int *out)
{
    out[0] = (int) offsetof(PageHeaderData, pd_flags);
    out[1] = (int) offsetof(PageHeaderData, pd_lower);
    out[2] = (int) offsetof(PageHeaderData, pd_upper);
    out[3] = (int) offsetof(PageHeaderData, pd_special);
    out[4] = (int) offsetof(PageHeaderData, pd_pagesize_version);
    out[5] = (int) offsetof(PageHeaderData, pd_linp);
    out[6] = (int) sizeof(ItemIdData);
    out[7] = (int) offsetof(HeapTupleHeaderData, t_ctid);
    out[8] = (int) offsetof(HeapTupleHeaderData, t_infomask2);
    out[9] = (int) offsetof(HeapTupleHeaderData, t_bits);
}
void ref_debug_array(void)
{
    double v[3] = {0, 0, 0};
    ArrayType *a = make_f8_array3(v);
    fprintf(stderr, "ndim %d dim0 %d hasnull %d elemtype %u FLOAT8OID %u sizeof(ArrayType) %zu overhead %zu\n",
            ARR_NDIM(a), ARR_DIMS(a)[0], ARR_HASNULL(a), ARR_ELEMTYPE(a), FLOAT8OID, sizeof(ArrayType), (size_t) ARR_OVERHEAD_NONULLS(1));
}
int ref_fcinfo_arg_offset(void) { return (int) offsetof(FunctionCallInfoData, arg); }

/* ---- forward-node pages (the redistribute wire format) ----------------------
 * Sender side as FragmentSendAttrs does it (executor/execFragment.c:2067-2136): header size and data size from
 * heap_minimal_tuple_header_size / heap_compute_data_size, the tuple formed in place by heap_form_minimal_tuple_ptr,
 * lower advanced by MAXALIGN(len); a new page when the aligned tuple does not fit (FragmentGetPage's rule,
 * execFragment.c:1858, restated here because that function is static and drags the whole buffer manager in);
 * pages initialised by FnPageInit (+ virtualid as FragmentGetPage sets it); the end of the stream is a MAX_UINT32
 * length word and FNPAGE_END (FragmentSendNullTuple, execFragment.c:1963-1975).  values/isnull are row-major. */
static Datum glue_datum(int16 attlen, int64 v)
{
    if (attlen == -1) {
        struct varlena *t = (struct varlena *) calloc(1, VARHDRSZ + 1);
        SET_VARSIZE(t, VARHDRSZ + 1);
        VARDATA(t)[0] = (char) v;
        return PointerGetDatum(t);
    }
    return (Datum) v;
}
int64 ref_fnpage_pack(int natts, const int16 *attlen, const int8 *attalign, const int64 *values, const uint8 *isnull, int64 nrows,
                      int64 qid_ts, int64 qid_seq, int fid, int nodeid, int workerid, int virtualid, int end_marker,
                      uint8 *pages, int64 cap_pages)
{
    TupleDesc d = make_desc(natts, attlen, attalign);
    Datum *v = (Datum *) calloc(natts, sizeof(Datum));
    bool *n = (bool *) calloc(natts, sizeof(bool));
    FNQueryId qid; qid.timestamp_nodeid = qid_ts; qid.sequence = qid_seq;
    int64 npages = 0;
    char *page = NULL;
#define GLUE_NEW_PAGE() do { if (npages >= cap_pages) { npages = -1; goto done; } page = (char *) pages + npages * BLCKSZ; npages++; \
        memset(page, 0, BLCKSZ); FnPageInit(page, qid, (uint16) fid, (uint16) nodeid, (uint16) workerid); \
        ((FnPageHeader) page)->virtualid = (uint8) virtualid; } while (0)
    for (int64 r = 0; r < nrows; r++) {
        bool hasnull = false;
        for (int i = 0; i < natts; i++) { n[i] = isnull[r * natts + i] != 0; hasnull |= n[i]; v[i] = n[i] ? (Datum) 0 : glue_datum(attlen[i], values[r * natts + i]); }
        int hoff = heap_minimal_tuple_header_size(d, hasnull);
        Size len = hoff + heap_compute_data_size(d, v, n);
        Size aligned = MAXALIGN(len);
        if (page == NULL || aligned > BLCKSZ - ((FnPageHeader) page)->lower) GLUE_NEW_PAGE();
        void *tuple = page + ((FnPageHeader) page)->lower;
        heap_form_minimal_tuple_ptr(d, v, n, hasnull, len, hoff, (MinimalTuple *) &tuple);
        ((FnPageHeader) page)->lower += MAXALIGN(len);
        for (int i = 0; i < natts; i++) if (attlen[i] == -1 && !n[i]) free(DatumGetPointer(v[i]));
    }
    if (end_marker) {
        uint32 n32 = MAX_UINT32;
        if (page == NULL || sizeof(uint32) > BLCKSZ - ((FnPageHeader) page)->lower) GLUE_NEW_PAGE();
        memcpy(page + ((FnPageHeader) page)->lower, &n32, sizeof(uint32));
        ((FnPageHeader) page)->lower += sizeof(uint32);
        FnPageSetFlag(page, FNPAGE_END);
    }
done:
    free(v); free(n); free(d);
    return npages;
}
/* Receiver side: the iterator macros of fnbufpage.h as the forward receiver uses them (executor/tqueueThread.c:913-925),
 * each item handed to heap_deform_tuple the way ExecStoreMinimalTuple presents a minimal tuple (t_data =
 * tuple - MINIMAL_TUPLE_OFFSET, executor/execTuples.c).  Returns the number of rows, -1 if cap_rows is too small,
 * -2 on a page carrying FNPAGE_HUGE. */
int64 ref_fnpage_unpack(const uint8 *pages, int64 npages, int natts, const int16 *attlen, const int8 *attalign,
                        int64 *values_out, uint8 *isnull_out, int64 cap_rows)
{
    TupleDesc d = make_desc(natts, attlen, attalign);
    Datum *v = (Datum *) calloc(natts, sizeof(Datum));
    bool *n = (bool *) calloc(natts, sizeof(bool));
    int64 rows = 0;
    for (int64 p = 0; p < npages && rows >= 0; p++) {
        char *page = (char *) pages + p * BLCKSZ;
        FnPageIterator iter;
        if (((FnPageHeader) page)->flag & FNPAGE_HUGE) { rows = -2; break; }
        InitFnPageIterator(&iter);
        while (!FnPageIterateDone(page, &iter)) {
            uint32 len; Pointer data;
            FnPageIterateNext(page, &iter, len, data);
            if (len == MAX_UINT32) break;
            if (rows >= cap_rows) { rows = -1; break; }
            HeapTupleData htup;
            memset(&htup, 0, sizeof(htup));
            htup.t_len = len + MINIMAL_TUPLE_OFFSET;
            htup.t_data = (HeapTupleHeader) ((char *) data - MINIMAL_TUPLE_OFFSET);
            heap_deform_tuple(&htup, d, v, n);
            for (int i = 0; i < natts; i++) {
                isnull_out[rows * natts + i] = n[i];
                if (n[i]) { values_out[rows * natts + i] = 0; continue; }
                if (attlen[i] == -1) values_out[rows * natts + i] = (int64) (int8) VARDATA_ANY((struct varlena *) DatumGetPointer(v[i]))[0];
                else if (attlen[i] == 1) values_out[rows * natts + i] = (int64) DatumGetChar(v[i]);
                else if (attlen[i] == 2) values_out[rows * natts + i] = (int64) DatumGetInt16(v[i]);
                else if (attlen[i] == 4) values_out[rows * natts + i] = (int64) DatumGetInt32(v[i]);
                else values_out[rows * natts + i] = DatumGetInt64(v[i]);
            }
            rows++;
        }
    }
    free(v); free(n); free(d);
    return rows;
}
int ref_sizeof_fnpage_header(void) { return (int) SizeOfFnPageHeaderData; }
int ref_invalid_shardid(void) { return (int) InvalidShardID; }
