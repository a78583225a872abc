/*
 * orc_exec.c — tuple-at-a-time restatement of the reference executor nodes on
 * the hot path:  SeqScan -> [HashJoin(Hash)] -> Agg(HASHED).
 * TEST INFRASTRUCTURE (see otb_oracle.h).  Also the CPU baseline timed by
 * bench.py (`cpu_baseline.kind = "port"`): it keeps the reference's cost
 * structure — one ExecProcNode indirect call per node per tuple, a visibility
 * pass per page, incremental slot_deform_tuple, a function-pointer call per
 * hash key / qual / transition, MinimalTuple copies into 32 KB dense chunks,
 * pointer-chased bucket chains, a robin-hood group table.
 *
 * Restated from (all under /root/reference/src/backend/executor unless noted):
 *   ExecScan / SeqNext            execScan.c:141-330, nodeSeqscan.c:60-124
 *   MultiExecPrivateHash          nodeHash.c:157-242
 *   ExecHashGetHashValue          nodeHash.c:2026-2112
 *   ExecHashTableInsert           nodeHash.c:1828-1910   dense_alloc :3002
 *   ExecHashIncreaseNumBuckets    nodeHash.c:1700-1790
 *   ExecHashGetBucketAndBatch     nodeHash.c:2142-2161
 *   ExecScanHashBucket            nodeHash.c:2174-2234
 *   ExecHashJoinImpl (INNER)      nodeHashjoin.c:186-742
 *   agg_fill_hash_table           nodeAgg.c:2609-2648
 *   lookup_hash_entries           nodeAgg.c:2149-2196
 *   LookupTupleHashEntry          execGrouping.c:295-321
 *   TupleHashTableHash_internal   execGrouping.c:415-473
 *   tuplehash_insert (robin hood) ../../include/lib/simplehash.h:540-700
 *   advance_transition_function   nodeAgg.c:742-840
 *   finalize_aggregates           nodeAgg.c:1363
 *   int8inc / int8pl              ../utils/adt/int8.c:714
 *   float8pl / float8mi / float8mul   ../utils/adt/float.c:970-1040
 *   float8_accum/_combine/_avg    ../utils/adt/float.c:2725-3008
 *   int4_sum                      ../utils/adt/numeric.c:6154-6204
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "orc_internal.h"

int orc_heapgetpage(const uint8_t *pg, uint16_t *vistuples);

/* ------------------------------------------------------------------ arena */
typedef struct arena_blk { struct arena_blk *next; size_t used, cap; } arena_blk;
typedef struct arena { arena_blk *head; } arena;
static void *arena_alloc(arena *a, size_t n)
{
    n = (n + 7) & ~(size_t) 7;
    if (!a->head || a->head->used + n > a->head->cap) {
        size_t cap = n > (1u << 20) ? n : (1u << 20);
        arena_blk *b = (arena_blk *) malloc(sizeof(arena_blk) + cap);
        b->next = a->head; b->used = 0; b->cap = cap; a->head = b;
    }
    void *p = (char *) (a->head + 1) + a->head->used;
    a->head->used += n;
    return p;
}
static void arena_free(arena *a)
{
    while (a->head) { arena_blk *n = a->head->next; free(a->head); a->head = n; }
}

/* ---------------------------------------------------------- error status */
static __thread int g_err;
static __thread double g_last_secs;
static __thread int64_t g_hs_nbuckets, g_hs_ntuples, g_hs_space;
double orc_last_exec_seconds(void) { return g_last_secs; }
static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* ------------------------------------------------------ Datum functions */
static inline double D2F(int64_t d) { double x; memcpy(&x, &d, 8); return x; }
static inline int64_t F2D(double x) { int64_t d; memcpy(&d, &x, 8); return d; }

/* float8_cmp_internal, float.c:1160-1195: NaN sorts after everything */
static int float8_cmp(double a, double b)
{
    if (isnan(a)) return isnan(b) ? 0 : 1;
    if (isnan(b)) return -1;
    return a > b ? 1 : (a < b ? -1 : 0);
}
typedef int (*cmp_fn)(int64_t, int64_t);
static int cmp_int(int64_t a, int64_t b) { return a > b ? 1 : (a < b ? -1 : 0); }
static int cmp_char(int64_t a, int64_t b)          /* charlt etc.: (uint8) compare, char.c */
{ uint8_t x = (uint8_t) a, y = (uint8_t) b; return x > y ? 1 : (x < y ? -1 : 0); }
static int cmp_f8(int64_t a, int64_t b) { return float8_cmp(D2F(a), D2F(b)); }
static cmp_fn cmp_for_type(int type)
{
    switch (type) {
        case GX_FLOAT8: return cmp_f8;
        case GX_CHAR: case ORC_BPCHAR1: return cmp_char;
        default: return cmp_int;
    }
}
static int op_holds(int op, int c)
{
    switch (op) {
        case GX_LT: return c < 0;  case GX_LE: return c <= 0; case GX_EQ: return c == 0;
        case GX_GE: return c >= 0; case GX_GT: return c > 0;  case GX_NE: return c != 0;
    }
    return 0;
}

/* CHECKFLOATVAL, float.c:40-60 */
static double checkfloat(double r, int inf_ok, int zero_ok)
{
    if (isinf(r) && !inf_ok) g_err = GX_ERR_OVERFLOW;
    if (r == 0.0 && !zero_ok) g_err = GX_ERR_OVERFLOW;    /* "underflow" */
    return r;
}
static double float8pl(double a, double b) { return checkfloat(a + b, isinf(a) || isinf(b), 1); }
static double float8mi(double a, double b) { return checkfloat(a - b, isinf(a) || isinf(b), 1); }
static double float8mul(double a, double b)
{ return checkfloat(a * b, isinf(a) || isinf(b), a == 0 || b == 0); }

/* ------------------------------------------------------------ PlanState */
typedef struct PlanState PlanState;
struct PlanState {
    orc_slot *(*ExecProcNode)(PlanState *);     /* execnodes.h:1036 */
    PlanState *lefttree, *righttree;
};
static inline orc_slot *ExecProcNode(PlanState *n) { return n->ExecProcNode(n); }
#define TupIsNull(s) ((s) == NULL || (s)->empty)

typedef struct QualItem { int col; int op; int64_t k; cmp_fn cmp; } QualItem;

/* ------------------------------------------------------------- SeqScan */
typedef struct SeqScanState {
    PlanState ps;
    const orc_rel *rel;
    int64_t   page;           /* rs_cblock */
    int       ntup, cindex;   /* rs_ntuples, rs_cindex */
    uint16_t  vis[ORC_BLCKSZ / 4];
    int       nquals;
    QualItem  quals[GX_MAX_PREDS];
    orc_slot  slot;           /* ss_ScanTupleSlot */
    orc_attr *attrs;          /* private copy: attcacheoff is per-descriptor */
} SeqScanState;

/* ExecQual over "col op const" clauses; NULL input makes a strict operator
 * return NULL, which fails the qual (execExprInterp.c EEOP_QUAL) */
static inline int ExecQual(int nquals, const QualItem *q, orc_slot *slot)
{
    for (int i = 0; i < nquals; i++) {
        uint8_t isnull;
        int64_t v = orc_slot_getattr(slot, q[i].col, &isnull);
        if (isnull) return 0;
        if (!op_holds(q[i].op, q[i].cmp(v, q[i].k))) return 0;
    }
    return 1;
}

/* SeqNext + heapgettup_pagemode, nodeSeqscan.c:60, heapam.c:921 */
static orc_slot *SeqNext(SeqScanState *node)
{
    for (;;) {
        if (node->cindex >= node->ntup) {
            node->page++;
            if (node->page >= node->rel->npages) { node->slot.empty = 1; return &node->slot; }
            node->ntup = orc_heapgetpage(node->rel->pages[node->page], node->vis);
            node->cindex = 0;
            continue;
        }
        const uint8_t *pg = node->rel->pages[node->page];
        uint32_t lp;
        memcpy(&lp, pg + ORC_PAGE_HDR + 4 * (uint32_t) (node->vis[node->cindex++] - 1), 4);
        /* ExecStoreBufferHeapTuple, execTuples.c:456 */
        node->slot.tuple = pg + (lp & 0x7FFF);
        node->slot.nvalid = 0; node->slot.off = 0; node->slot.slow = 0; node->slot.empty = 0;
        return &node->slot;
    }
}
/* ExecScan, execScan.c:141-330 (no projection: physical tlist) */
static orc_slot *ExecSeqScan(PlanState *ps)
{
    SeqScanState *node = (SeqScanState *) ps;
    for (;;) {
        orc_slot *slot = SeqNext(node);
        if (TupIsNull(slot)) return slot;
        if (node->nquals == 0 || ExecQual(node->nquals, node->quals, slot))
            return slot;
    }
}
static SeqScanState *ExecInitSeqScan(const orc_rel *rel, int npreds, const gx_pred *preds)
{
    SeqScanState *n = (SeqScanState *) calloc(1, sizeof(*n));
    n->ps.ExecProcNode = ExecSeqScan;
    n->rel = rel;
    n->page = -1;
    n->attrs = (orc_attr *) malloc(sizeof(orc_attr) * (size_t) rel->natts);
    memcpy(n->attrs, rel->attrs, sizeof(orc_attr) * (size_t) rel->natts);
    for (int i = 0; i < rel->natts; i++) n->attrs[i].attcacheoff = -1;
    n->slot.natts = rel->natts;
    n->slot.attrs = n->attrs;
    n->slot.values = (int64_t *) calloc((size_t) rel->natts, 8);
    n->slot.isnull = (uint8_t *) calloc((size_t) rel->natts, 1);
    n->slot.empty = 1;
    n->nquals = npreds;
    for (int i = 0; i < npreds; i++) {
        int t = rel->attrs[preds[i].col].type;
        n->quals[i].col = preds[i].col;
        n->quals[i].op = preds[i].op;
        n->quals[i].k = (t == GX_FLOAT8) ? F2D(preds[i].fval) : preds[i].ival;
        n->quals[i].cmp = cmp_for_type(t);
    }
    return n;
}
static void ExecEndSeqScan(SeqScanState *n)
{
    free(n->attrs); free(n->slot.values); free(n->slot.isnull); free(n);
}

/* ---------------------------------------------------------------- Hash */
#define HASH_CHUNK_SIZE   (32 * 1024)
#define HJTUPLE_OVERHEAD  24          /* MAXALIGN(sizeof(HashJoinTupleData)), hashjoin.h:79-91 */
#define MINTUP_HDR        16          /* data offset of a MinimalTuple without nulls */

typedef struct HashJoinTupleData {
    struct HashJoinTupleData *next_unshared;
    uint64_t next_shared;             /* dsa_pointer half of HashJoinTupleNext */
    uint32_t hashvalue;
    /* MinimalTuple follows at HJTUPLE_OVERHEAD */
} HashJoinTupleData;
typedef HashJoinTupleData *HashJoinTuple;

typedef struct HashChunk { struct HashChunk *next; size_t used, maxlen; int ntuples; } HashChunk;
#define HASH_CHUNK_HEADER_SIZE ((sizeof(HashChunk) + 7) & ~(size_t) 7)

typedef struct HashJoinTable {
    int64_t nbuckets, nbuckets_optimal;
    int     log2_nbuckets;
    HashJoinTuple *buckets;
    HashChunk *chunks;
    double  totalTuples;
    size_t  spaceUsed;
    int     nkeys;                    /* one key column in this build */
    int     ninner;                   /* attrs in stored minimal tuples */
    orc_attr inner_attrs[1 + GX_MAX_PAYLOAD];
} HashJoinTable;

/* dense_alloc, nodeHash.c:3002-3070 */
static void *dense_alloc(HashJoinTable *ht, size_t size)
{
    size = (size + 7) & ~(size_t) 7;
    if (!ht->chunks || ht->chunks->maxlen - ht->chunks->used < size) {
        HashChunk *c = (HashChunk *) malloc(HASH_CHUNK_HEADER_SIZE + HASH_CHUNK_SIZE);
        c->maxlen = HASH_CHUNK_SIZE; c->used = 0; c->ntuples = 0;
        c->next = ht->chunks; ht->chunks = c;
    }
    void *p = (char *) ht->chunks + HASH_CHUNK_HEADER_SIZE + ht->chunks->used;
    ht->chunks->used += size; ht->chunks->ntuples++;
    return p;
}

typedef uint32_t (*hash_fn)(int type, int64_t datum);

/* ExecHashGetHashValue for a single key: rotate, then XOR the CRC32C "new"
 * hash (HashFuncAssign).  Returns 0 when the key is NULL (hashStrict). */
static inline int ExecHashGetHashValue(int type, int64_t key, int isnull, uint32_t *hashvalue)
{
    uint32_t hashkey = 0;
    hashkey = (hashkey << 1) | (hashkey >> 31);       /* pg_rotate_left32(hashkey, 1) */
    if (isnull) return 0;                             /* cannot match */
    hashkey ^= orc_hash_datum_new(type, key);
    *hashvalue = hashkey;
    return 1;
}

/* ExecHashIncreaseNumBuckets, nodeHash.c:1700-1790: relink every tuple */
static void ExecHashIncreaseNumBuckets(HashJoinTable *ht)
{
    if (ht->nbuckets >= ht->nbuckets_optimal) return;
    ht->nbuckets = ht->nbuckets_optimal;
    ht->log2_nbuckets = 0;
    while (((int64_t) 1 << ht->log2_nbuckets) < ht->nbuckets) ht->log2_nbuckets++;
    free(ht->buckets);
    ht->buckets = (HashJoinTuple *) calloc((size_t) ht->nbuckets, sizeof(HashJoinTuple));
    for (HashChunk *c = ht->chunks; c; c = c->next) {
        size_t idx = 0;
        while (idx < c->used) {
            HashJoinTuple t = (HashJoinTuple) ((char *) c + HASH_CHUNK_HEADER_SIZE + idx);
            uint32_t tlen; memcpy(&tlen, (char *) t + HJTUPLE_OVERHEAD, 4);
            uint32_t b = t->hashvalue & (uint32_t) (ht->nbuckets - 1);
            t->next_unshared = ht->buckets[b];
            ht->buckets[b] = t;
            idx += (HJTUPLE_OVERHEAD + tlen + 7) & ~(size_t) 7;
        }
    }
}

static void ExecHashTableDestroy(HashJoinTable *ht)
{
    while (ht->chunks) { HashChunk *n = ht->chunks->next; free(ht->chunks); ht->chunks = n; }
    free(ht->buckets); free(ht);
}

/* MultiExecPrivateHash, nodeHash.c:157-242 */
static HashJoinTable *MultiExecHash(PlanState *inner_scan, const orc_rel *inner,
                                    const orc_join_spec *js)
{
    HashJoinTable *ht = (HashJoinTable *) calloc(1, sizeof(*ht));
    int keytype = inner->attrs[js->inner_key_col].type;
    ht->nbuckets = ht->nbuckets_optimal = 1024;      /* ExecChooseHashTableSize floor */
    ht->log2_nbuckets = 10;
    ht->buckets = (HashJoinTuple *) calloc(1024, sizeof(HashJoinTuple));
    ht->ninner = 1 + js->n_payload;
    ht->inner_attrs[0] = inner->attrs[js->inner_key_col];
    for (int i = 0; i < js->n_payload; i++) ht->inner_attrs[1 + i] = inner->attrs[js->payload_cols[i]];
    for (int i = 0; i < ht->ninner; i++) ht->inner_attrs[i].attcacheoff = -1;

    for (;;) {
        orc_slot *slot = ExecProcNode(inner_scan);
        if (TupIsNull(slot)) break;
        uint8_t knull;
        int64_t key = orc_slot_getattr(slot, js->inner_key_col, &knull);
        uint32_t hashvalue;
        if (!ExecHashGetHashValue(keytype, key, knull, &hashvalue))
            continue;                                  /* NULL key: drop (strict) */
        /* ExecFetchSlotMinimalTuple: heap_form_minimal_tuple of the projected row */
        int64_t vals[1 + GX_MAX_PAYLOAD]; uint8_t nulls[1 + GX_MAX_PAYLOAD]; int hasnull = 0;
        vals[0] = key; nulls[0] = 0;
        for (int i = 0; i < js->n_payload; i++) {
            vals[1 + i] = orc_slot_getattr(slot, js->payload_cols[i], &nulls[1 + i]);
            hasnull |= nulls[1 + i];
        }
        uint32_t hoff = 15 + (hasnull ? (uint32_t) ((ht->ninner + 7) / 8) : 0);
        hoff = (hoff + 7) & ~7u;                       /* minimal-tuple data offset */
        uint32_t dlen = orc_compute_data_size(ht->inner_attrs, ht->ninner, vals, nulls);
        uint32_t tlen = hoff + dlen;
        /* ExecHashTableInsert, nodeHash.c:1828 */
        size_t hsz = HJTUPLE_OVERHEAD + tlen;
        HashJoinTuple ht_tup = (HashJoinTuple) dense_alloc(ht, hsz);
        ht_tup->hashvalue = hashvalue;
        uint8_t *mt = (uint8_t *) ht_tup + HJTUPLE_OVERHEAD;
        memset(mt, 0, hoff);
        memcpy(mt, &tlen, 4);                          /* t_len */
        { uint16_t im2 = (uint16_t) ht->ninner; memcpy(mt + 6, &im2, 2); }
        { uint16_t im = hasnull ? HEAP_HASNULL : 0; memcpy(mt + 8, &im, 2); }
        mt[14] = (uint8_t) (hoff + ORC_MINIMAL_TUPLE_OFFSET);   /* t_hoff as a full header */
        orc_form_data(ht->inner_attrs, ht->ninner, vals, nulls, hasnull ? mt + 15 : NULL, mt + hoff);
        uint32_t b = hashvalue & (uint32_t) (ht->nbuckets - 1);
        ht_tup->next_unshared = ht->buckets[b];
        ht->buckets[b] = ht_tup;
        ht->totalTuples += 1;
        if (ht->totalTuples > (double) ht->nbuckets_optimal)     /* NTUP_PER_BUCKET 1 */
            ht->nbuckets_optimal *= 2;
        ht->spaceUsed += hsz;
    }
    ExecHashIncreaseNumBuckets(ht);     /* "resize the hash table if needed", nodeHash.c:230 */
    g_hs_nbuckets = ht->nbuckets; g_hs_ntuples = (int64_t) ht->totalTuples;
    g_hs_space = (int64_t) (ht->spaceUsed + (size_t) ht->nbuckets * sizeof(HashJoinTuple));
    return ht;
}

/* ------------------------------------------------------------ HashJoin */
enum { HJ_BUILD_HASHTABLE = 1, HJ_NEED_NEW_OUTER, HJ_SCAN_BUCKET, HJ_FILL_OUTER_TUPLE };   /* nodeHashjoin.c:139-144 */
enum { ORC_JOIN_INNER = 0, ORC_JOIN_LEFT = 1, ORC_JOIN_SEMI = 4, ORC_JOIN_ANTI = 5 };   /* JoinType, nodes/nodes.h */

typedef struct HashJoinState {
    PlanState ps;
    int       state;
    PlanState *outer, *inner;
    const orc_rel *inner_rel;
    const orc_join_spec *js;
    int       outer_key_col, outer_key_type;
    HashJoinTable *ht;
    orc_slot *outer_slot;             /* hj_OuterTupleSlot */
    uint32_t  cur_hash;               /* hj_CurHashValue */
    HashJoinTuple cur_tuple;          /* hj_CurTuple */
    int64_t   cur_key;
    orc_slot  hashtup_slot;           /* hj_HashTupleSlot (minimal tuple) */
    int       single_match;           /* inner_unique || JOIN_SEMI, nodeHashjoin.c:859-861 */
    int       jointype;
    int       matched_outer;          /* hj_MatchedOuter */
    cmp_fn    key_cmp;
    /* projection: result slot = [needed outer attrs..., inner key, payload...] */
    int       n_outer_atts;
    orc_slot  result;                 /* ps_ResultTupleSlot, virtual */
    uint8_t  *outer_needed;           /* which outer attrs the parent reads */
} HashJoinState;

/* ExecScanHashBucket, nodeHash.c:2174-2234 */
static int ExecScanHashBucket(HashJoinState *hj)
{
    HashJoinTuple t = hj->cur_tuple;
    if (t) t = t->next_unshared;
    else t = hj->ht->buckets[hj->cur_hash & (uint32_t) (hj->ht->nbuckets - 1)];
    while (t) {
        if (t->hashvalue == hj->cur_hash) {
            /* ExecStoreMinimalTuple + ExecQualAndReset(hashclauses) */
            orc_slot *s = &hj->hashtup_slot;
            s->tuple = (const uint8_t *) t + HJTUPLE_OVERHEAD - ORC_MINIMAL_TUPLE_OFFSET;
            s->nvalid = 0; s->off = 0; s->slow = 0; s->empty = 0;
            uint8_t n;
            int64_t ik = orc_slot_getattr(s, 0, &n);
            if (!n && hj->key_cmp(ik, hj->cur_key) == 0) { hj->cur_tuple = t; return 1; }
        }
        t = t->next_unshared;
    }
    return 0;
}

/* ExecHashJoinImpl, nodeHashjoin.c:186-742: JOIN_INNER, JOIN_LEFT, JOIN_SEMI, JOIN_ANTI arms
 * (HJ_FILL_INNER_TUPLES — right/full joins — is not restated) */
static orc_slot *ExecHashJoin(PlanState *ps)
{
    HashJoinState *hj = (HashJoinState *) ps;
    for (;;) {
        switch (hj->state) {
            case HJ_BUILD_HASHTABLE:
                hj->ht = MultiExecHash(hj->inner, hj->inner_rel, hj->js);
                hj->hashtup_slot.natts = hj->ht->ninner;
                hj->hashtup_slot.attrs = hj->ht->inner_attrs;
                hj->state = HJ_NEED_NEW_OUTER;
                /* FALLTHROUGH */
            case HJ_NEED_NEW_OUTER: {
                /* ExecHashJoinOuterGetTuple, nodeHashjoin.c:1054 */
                orc_slot *o = ExecProcNode(hj->outer);
                if (TupIsNull(o)) { hj->result.empty = 1; return &hj->result; }
                uint8_t knull;
                int64_t key = orc_slot_getattr(o, hj->outer_key_col, &knull);
                hj->outer_slot = o; hj->matched_outer = 0;
                if (!ExecHashGetHashValue(hj->outer_key_type, key, knull, &hj->cur_hash)) {
                    /* NULL outer key never matches: ExecHashJoinOuterGetTuple skips the tuple unless the join
                     * fills outer tuples (HJ_FILL_OUTER, nodeHashjoin.c:1079-1090: keep_nulls) */
                    if (hj->jointype == ORC_JOIN_LEFT || hj->jointype == ORC_JOIN_ANTI) { hj->state = HJ_FILL_OUTER_TUPLE; continue; }
                    continue;
                }
                hj->cur_key = key; hj->cur_tuple = NULL;
                hj->state = HJ_SCAN_BUCKET;
            }   /* FALLTHROUGH */
            case HJ_SCAN_BUCKET: {
                if (!ExecScanHashBucket(hj)) { hj->state = HJ_FILL_OUTER_TUPLE; continue; }      /* :543-557 */
                hj->matched_outer = 1;                                                             /* :578 */
                if (hj->jointype == ORC_JOIN_ANTI) { hj->state = HJ_NEED_NEW_OUTER; continue; }    /* :628-634 */
                if (hj->single_match) hj->state = HJ_NEED_NEW_OUTER;                              /* :640-643 */
                /* ExecProject into the virtual result slot */
                orc_slot *r = &hj->result, *o = hj->outer_slot, *in = &hj->hashtup_slot;
                for (int a = 0; a < hj->n_outer_atts; a++)
                    if (hj->outer_needed[a])
                        r->values[a] = orc_slot_getattr(o, a, &r->isnull[a]);
                for (int a = 0; a < hj->ht->ninner; a++)
                    r->values[hj->n_outer_atts + a] = orc_slot_getattr(in, a, &r->isnull[hj->n_outer_atts + a]);
                r->empty = 0;
                return r;
            }
            case HJ_FILL_OUTER_TUPLE: {                                                            /* :668-689 */
                hj->state = HJ_NEED_NEW_OUTER;
                if (hj->matched_outer || !(hj->jointype == ORC_JOIN_LEFT || hj->jointype == ORC_JOIN_ANTI)) continue;
                /* the outer tuple joined with hj_NullInnerTupleSlot */
                orc_slot *r = &hj->result, *o = hj->outer_slot;
                for (int a = 0; a < hj->n_outer_atts; a++)
                    if (hj->outer_needed[a])
                        r->values[a] = orc_slot_getattr(o, a, &r->isnull[a]);
                for (int a = 0; a < hj->ht->ninner; a++) { r->values[hj->n_outer_atts + a] = 0; r->isnull[hj->n_outer_atts + a] = 1; }
                r->empty = 0;
                return r;
            }
        }
    }
}

static HashJoinState *ExecInitHashJoin(PlanState *outer, const orc_rel *outer_rel, int outer_key_col,
                                       PlanState *inner, const orc_rel *inner_rel,
                                       const orc_join_spec *js)
{
    HashJoinState *hj = (HashJoinState *) calloc(1, sizeof(*hj));
    hj->ps.ExecProcNode = ExecHashJoin;
    hj->state = HJ_BUILD_HASHTABLE;
    hj->outer = outer; hj->inner = inner; hj->inner_rel = inner_rel; hj->js = js;
    hj->outer_key_col = outer_key_col;
    hj->outer_key_type = outer_rel->attrs[outer_key_col].type;
    hj->key_cmp = cmp_for_type(hj->outer_key_type);
    hj->jointype = js->jointype;
    hj->single_match = js->inner_unique || js->jointype == ORC_JOIN_SEMI;      /* nodeHashjoin.c:859-861 */
    hj->n_outer_atts = outer_rel->natts;
    int nres = outer_rel->natts + 1 + js->n_payload;
    hj->result.natts = nres; hj->result.nvalid = nres; hj->result.tuple = NULL;
    hj->result.values = (int64_t *) calloc((size_t) nres, 8);
    hj->result.isnull = (uint8_t *) calloc((size_t) nres, 1);
    hj->result.empty = 1;
    hj->hashtup_slot.values = (int64_t *) calloc(1 + GX_MAX_PAYLOAD, 8);
    hj->hashtup_slot.isnull = (uint8_t *) calloc(1 + GX_MAX_PAYLOAD, 1);
    hj->outer_needed = (uint8_t *) calloc((size_t) outer_rel->natts, 1);
    return hj;
}
static void ExecEndHashJoin(HashJoinState *hj)
{
    if (hj->ht) ExecHashTableDestroy(hj->ht);
    free(hj->result.values); free(hj->result.isnull);
    free(hj->hashtup_slot.values); free(hj->hashtup_slot.isnull);
    free(hj->outer_needed); free(hj);
}

/* ------------------------------------------------------------------ Agg */
typedef struct AggStatePerGroupData {       /* nodeAgg.h:248-265 */
    int64_t transValue;
    uint8_t transValueIsNull;
    uint8_t noTransValue;
} AggStatePerGroupData;

typedef struct TupleHashEntryData {         /* execnodes.h:763-773 */
    uint8_t *firstTuple;                    /* MinimalTuple of the group columns */
    AggStatePerGroupData *additional;
    uint32_t status;
    uint32_t hash;
} TupleHashEntryData;

typedef struct tuplehash {
    uint64_t size; uint32_t members, sizemask, grow_threshold;
    TupleHashEntryData *data;
} tuplehash;

#define SH_FILLFACTOR 0.9
#define SH_GROW_MAX_DIB 25
#define SH_GROW_MAX_MOVE 150
#define SH_GROW_MIN_FILLFACTOR 0.1

static void sh_compute_parameters(tuplehash *tb, uint64_t newsize)
{
    uint64_t size = newsize < 2 ? 2 : newsize, p = 1;
    while (p < size) p <<= 1;
    tb->size = p; tb->sizemask = (uint32_t) (p - 1);
    tb->grow_threshold = (uint32_t) ((double) p * SH_FILLFACTOR);
}
static void sh_grow(tuplehash *tb, uint64_t newsize)
{
    uint64_t oldsize = tb->size;
    TupleHashEntryData *olddata = tb->data, *newdata;
    uint32_t startelem = 0, copyelem;
    sh_compute_parameters(tb, newsize);
    newdata = tb->data = (TupleHashEntryData *) calloc(tb->size, sizeof(TupleHashEntryData));
    for (uint32_t i = 0; i < oldsize; i++) {
        TupleHashEntryData *e = &olddata[i];
        if (e->status != 1) { startelem = i; break; }
        if ((e->hash & tb->sizemask) == i) { startelem = i; break; }
    }
    copyelem = startelem;
    for (uint32_t i = 0; i < oldsize; i++) {
        TupleHashEntryData *e = &olddata[copyelem];
        if (e->status == 1) {
            uint32_t cur = e->hash & tb->sizemask;
            while (newdata[cur].status != 0) cur = (cur + 1) & tb->sizemask;
            newdata[cur] = *e;
        }
        copyelem++;
        if (copyelem >= oldsize) copyelem = 0;
    }
    free(olddata);
}

typedef struct AggTrans {
    int fn;
    int argcol, argtype;              /* single-column argument, or -1 */
    gx_expr expr;
} AggTrans;

typedef struct AggState {
    PlanState *child;
    int ngroupcols, naggs;
    int grp_att[GX_MAX_GROUP_COLS];   /* attribute number in the child's slot */
    orc_attr grp_attrs[GX_MAX_GROUP_COLS];
    cmp_fn grp_eq[GX_MAX_GROUP_COLS];
    AggTrans trans[GX_MAX_AGGS];
    tuplehash tb;
    arena tablecxt;
    uint32_t hash_iv;
    orc_slot keyslot;                 /* tableslot for stored minimal tuples */
    const orc_attr *child_attrs;      /* type of each child attribute */
} AggState;

/* TupleHashTableHash_internal, execGrouping.c:415-473 (Jenkins per column) */
static uint32_t TupleHashTableHash(AggState *st, const int64_t *keys, const uint8_t *nulls)
{
    uint32_t hashkey = st->hash_iv;
    for (int i = 0; i < st->ngroupcols; i++) {
        hashkey = (hashkey << 1) | (hashkey >> 31);
        if (!nulls[i])
            hashkey ^= orc_hash_datum(st->grp_attrs[i].type == ORC_BPCHAR1 ? GX_CHAR : st->grp_attrs[i].type, keys[i]);
    }
    return orc_murmurhash32(hashkey);
}

/* TupleHashTableMatch, execGrouping.c:524: NULLs group together (not distinct) */
static int TupleHashTableMatch(AggState *st, const TupleHashEntryData *e, const int64_t *keys,
                               const uint8_t *nulls)
{
    orc_slot *s = &st->keyslot;
    s->tuple = e->firstTuple - ORC_MINIMAL_TUPLE_OFFSET;
    s->nvalid = 0; s->off = 0; s->slow = 0;
    for (int i = 0; i < st->ngroupcols; i++) {
        uint8_t n; int64_t v = orc_slot_getattr(s, i, &n);
        if (n != nulls[i]) return 0;
        if (!n && st->grp_eq[i](v, keys[i]) != 0) return 0;
    }
    return 1;
}

/* initialize_aggregate, nodeAgg.c:614-700: initcond per pg_aggregate.h */
static void initialize_hash_entry(AggState *st, TupleHashEntryData *e)
{
    e->additional = (AggStatePerGroupData *) arena_alloc(&st->tablecxt, sizeof(AggStatePerGroupData) * (size_t) (st->naggs ? st->naggs : 1));
    for (int a = 0; a < st->naggs; a++) {
        AggStatePerGroupData *pg = &e->additional[a];
        switch (st->trans[a].fn) {
            case GX_AGG_COUNT_STAR: case GX_AGG_COUNT:           /* agginitval "0" */
                pg->transValue = 0; pg->transValueIsNull = 0; pg->noTransValue = 0; break;
            case GX_AGG_AVG_F8: {                                /* agginitval "{0,0,0}" */
                double *tv = (double *) arena_alloc(&st->tablecxt, 3 * sizeof(double));
                tv[0] = tv[1] = tv[2] = 0.0;
                pg->transValue = (int64_t) (intptr_t) tv; pg->transValueIsNull = 0; pg->noTransValue = 0; break;
            }
            default:                                             /* agginitval NULL */
                pg->transValue = 0; pg->transValueIsNull = 1; pg->noTransValue = 1; break;
        }
    }
}

/* tuplehash_insert, simplehash.h:540-700 */
static TupleHashEntryData *LookupTupleHashEntry(AggState *st, const int64_t *keys, const uint8_t *nulls)
{
    tuplehash *tb = &st->tb;
    uint32_t hash = TupleHashTableHash(st, keys, nulls);
    uint32_t startelem, curelem, insertdist;
    TupleHashEntryData *data, *entry;
restart:
    insertdist = 0;
    if (tb->members >= tb->grow_threshold)
        sh_grow(tb, tb->size * 2);
    data = tb->data;
    startelem = hash & tb->sizemask;
    curelem = startelem;
    for (;;) {
        entry = &data[curelem];
        if (entry->status == 0) {
            tb->members++;
            goto fill;
        }
        if (entry->hash == hash && TupleHashTableMatch(st, entry, keys, nulls))
            return entry;
        {
            uint32_t curoptimal = entry->hash & tb->sizemask;
            uint32_t curdist = (curoptimal <= curelem) ? curelem - curoptimal
                                                       : (uint32_t) (tb->size + curelem - curoptimal);
            if (insertdist > curdist) {
                TupleHashEntryData *lastentry = entry;
                uint32_t emptyelem = curelem, moveelem;
                int32_t emptydist = 0;
                for (;;) {
                    TupleHashEntryData *emptyentry;
                    emptyelem = (emptyelem + 1) & tb->sizemask;
                    emptyentry = &data[emptyelem];
                    if (emptyentry->status == 0) { lastentry = emptyentry; break; }
                    if (++emptydist > SH_GROW_MAX_MOVE &&
                        ((double) tb->members / (double) tb->size) >= SH_GROW_MIN_FILLFACTOR) {
                        tb->grow_threshold = 0;
                        goto restart;
                    }
                }
                moveelem = emptyelem;
                while (moveelem != curelem) {
                    TupleHashEntryData *moveentry;
                    moveelem = (moveelem - 1) & tb->sizemask;
                    moveentry = &data[moveelem];
                    *lastentry = *moveentry;
                    lastentry = moveentry;
                }
                tb->members++;
                goto fill;
            }
        }
        curelem = (curelem + 1) & tb->sizemask;
        insertdist++;
        if (insertdist > SH_GROW_MAX_DIB &&
            ((double) tb->members / (double) tb->size) >= SH_GROW_MIN_FILLFACTOR) {
            tb->grow_threshold = 0;
            goto restart;
        }
    }
fill:
    entry->status = 1;
    entry->hash = hash;
    {   /* ExecCopySlotMinimalTuple of the group columns into tablecxt */
        int hasnull = 0;
        for (int i = 0; i < st->ngroupcols; i++) hasnull |= nulls[i];
        uint32_t hoff = 15 + (hasnull ? (uint32_t) ((st->ngroupcols + 7) / 8) : 0);
        hoff = (hoff + 7) & ~7u;
        uint32_t dlen = orc_compute_data_size(st->grp_attrs, st->ngroupcols, keys, nulls);
        uint32_t tlen = hoff + dlen;
        uint8_t *mt = (uint8_t *) arena_alloc(&st->tablecxt, tlen ? tlen : 8);
        memset(mt, 0, hoff);
        memcpy(mt, &tlen, 4);
        { uint16_t im2 = (uint16_t) st->ngroupcols; memcpy(mt + 6, &im2, 2); }
        { uint16_t im = hasnull ? HEAP_HASNULL : 0; memcpy(mt + 8, &im, 2); }
        mt[14] = (uint8_t) (hoff + ORC_MINIMAL_TUPLE_OFFSET);
        orc_form_data(st->grp_attrs, st->ngroupcols, keys, nulls, hasnull ? mt + 15 : NULL, mt + hoff);
        entry->firstTuple = mt;
    }
    initialize_hash_entry(st, entry);
    return entry;
}

/* the aggregate's argument expression (ExecEvalExpr over float8 operators:
 * float8pl/float8mi/float8mul are strict — any NULL input yields NULL) */
static double eval_expr(const gx_expr *e, orc_slot *slot, const orc_attr *attrs, uint8_t *isnull)
{
    double st[GX_MAX_EXPR_OPS]; uint8_t sn[GX_MAX_EXPR_OPS]; int sp = 0;
    for (int i = 0; i < e->nops; i++) {
        const gx_expr_op *op = &e->ops[i];
        switch (op->op) {
            case GX_OP_COL: {
                uint8_t n; int64_t v = orc_slot_getattr(slot, op->col, &n);
                sn[sp] = n;
                st[sp] = n ? 0.0 : (attrs[op->col].type == GX_FLOAT8 ? D2F(v) : (double) v);  /* i4tod/i8tod */
                sp++; break;
            }
            case GX_OP_CONST: st[sp] = op->k; sn[sp] = 0; sp++; break;
            default: {
                double b = st[--sp], a = st[--sp]; uint8_t n = sn[sp] | sn[sp + 1];
                double r = 0.0;
                if (!n) r = op->op == GX_OP_ADD ? float8pl(a, b) : op->op == GX_OP_SUB ? float8mi(a, b) : float8mul(a, b);
                st[sp] = r; sn[sp] = n; sp++;
            }
        }
    }
    *isnull = sn[0];
    return st[0];
}

/* advance_aggregates -> advance_transition_function, nodeAgg.c:742-863 */
static void advance_aggregates(AggState *st, AggStatePerGroupData *pergroup, orc_slot *slot)
{
    for (int a = 0; a < st->naggs; a++) {
        AggTrans *t = &st->trans[a];
        AggStatePerGroupData *pg = &pergroup[a];
        switch (t->fn) {
            case GX_AGG_COUNT_STAR:                 /* int8inc, int8.c:714 (overflow-checked) */
                if (pg->transValue == INT64_MAX) g_err = GX_ERR_OVERFLOW; else pg->transValue++;
                break;
            case GX_AGG_COUNT: {                    /* int8inc_any: strict on the input */
                uint8_t n; (void) orc_slot_getattr(slot, t->argcol, &n);
                if (!n) pg->transValue++;
                break;
            }
            case GX_AGG_SUM_F8: case GX_AGG_MIN_F8: case GX_AGG_MAX_F8: {
                uint8_t n; double x = eval_expr(&t->expr, slot, st->child_attrs, &n);
                if (n) break;                       /* strict transfn: NULL input ignored */
                if (pg->noTransValue) {             /* first non-NULL input becomes the state */
                    pg->transValue = F2D(x); pg->transValueIsNull = 0; pg->noTransValue = 0; break;
                }
                if (pg->transValueIsNull) break;
                double s = D2F(pg->transValue);
                if (t->fn == GX_AGG_SUM_F8) s = float8pl(s, x);
                else if (t->fn == GX_AGG_MIN_F8) s = (float8_cmp(s, x) < 0) ? s : x;   /* float8smaller */
                else s = (float8_cmp(s, x) > 0) ? s : x;                               /* float8larger */
                pg->transValue = F2D(s);
                break;
            }
            case GX_AGG_AVG_F8: {                   /* float8_accum, float.c:2823-2903 */
                uint8_t n; double newval = eval_expr(&t->expr, slot, st->child_attrs, &n);
                if (n) break;
                double *tv = (double *) (intptr_t) pg->transValue;
                double N = tv[0], Sx = tv[1], Sxx = tv[2], tmp;
                N += 1.0; Sx += newval;
                if (tv[0] > 0.0) {
                    tmp = newval * N - Sx;
                    Sxx += tmp * tmp / (N * tv[0]);
                    if (isinf(Sx) || isinf(Sxx)) {
                        if (!isinf(tv[1]) && !isinf(newval)) g_err = GX_ERR_OVERFLOW;
                        Sxx = NAN;
                    }
                } else if (isnan(newval) || isinf(newval)) Sxx = NAN;
                tv[0] = N; tv[1] = Sx; tv[2] = Sxx;
                break;
            }
            case GX_AGG_SUM_I4: case GX_AGG_SUM_I8: {  /* int4_sum (non-strict, NULL init) */
                uint8_t n; int64_t v = orc_slot_getattr(slot, t->argcol, &n);
                if (n) break;
                if (pg->transValueIsNull) { pg->transValue = v; pg->transValueIsNull = 0; pg->noTransValue = 0; }
                else if (__builtin_add_overflow(pg->transValue, v, &pg->transValue)) g_err = GX_ERR_OVERFLOW;
                break;
            }
        }
    }
}

static void fill_result(AggState *st, orc_result *out)
{
    tuplehash *tb = &st->tb;
    int ng = st->ngroupcols, na = st->naggs;
    int64_t n = tb->members;
    out->n_group_cols = ng; out->n_aggs = na; out->ngroups = n;
    out->keys = (int64_t *) calloc((size_t) (n * (ng ? ng : 1)), 8);
    out->aggs = (double *) calloc((size_t) (n * (na ? na : 1)), 8);
    out->nulls = (uint8_t *) calloc((size_t) (n * (ng + na ? ng + na : 1)), 1);
    out->states = (double *) calloc((size_t) (n * (na ? na : 1) * 3), 8);
    int64_t g = 0;
    /* agg_retrieve_hash_table_in_memory, nodeAgg.c:2840: table order */
    for (uint64_t i = 0; i < tb->size; i++) {
        TupleHashEntryData *e = &tb->data[i];
        if (e->status != 1) continue;
        orc_slot *s = &st->keyslot;
        s->tuple = e->firstTuple - ORC_MINIMAL_TUPLE_OFFSET; s->nvalid = 0; s->off = 0; s->slow = 0;
        for (int c = 0; c < ng; c++) {
            uint8_t nl; int64_t v = orc_slot_getattr(s, c, &nl);
            out->keys[g * ng + c] = nl ? 0 : v;
            out->nulls[g * (ng + na) + c] = nl;
        }
        for (int a = 0; a < na; a++) {           /* finalize_aggregates, nodeAgg.c:1363 */
            AggStatePerGroupData *pg = &e->additional[a];
            double res = 0.0; uint8_t nl = 0; double *stt = &out->states[(g * na + a) * 3];
            switch (st->trans[a].fn) {
                case GX_AGG_COUNT_STAR: case GX_AGG_COUNT:
                    memcpy(&res, &pg->transValue, 8); memcpy(&stt[0], &pg->transValue, 8); break;
                case GX_AGG_SUM_I4: case GX_AGG_SUM_I8:
                    nl = pg->transValueIsNull; memcpy(&res, &pg->transValue, 8);
                    memcpy(&stt[0], &pg->transValue, 8); stt[1] = nl ? 1.0 : 0.0; break;
                case GX_AGG_AVG_F8: {            /* float8_avg, float.c:2991-3008 */
                    double *tv = (double *) (intptr_t) pg->transValue;
                    stt[0] = tv[0]; stt[1] = tv[1]; stt[2] = tv[2];
                    if (tv[0] == 0.0) nl = 1; else res = tv[1] / tv[0];
                    break;
                }
                default:
                    nl = pg->transValueIsNull; res = D2F(pg->transValue);
                    stt[0] = nl ? 0.0 : 1.0; stt[1] = res; break;
            }
            out->aggs[g * na + a] = nl ? 0.0 : res;
            out->nulls[g * (ng + na) + ng + a] = nl;
        }
        g++;
    }
}

void orc_result_free(orc_result *r)
{
    free(r->keys); free(r->aggs); free(r->nulls); free(r->states);
    memset(r, 0, sizeof(*r));
}

int orc_exec_agg(const orc_rel *outer, const orc_rel *inner, const orc_join_spec *join,
                 const gx_agg_plan *plan, orc_result *out)
{
    g_err = 0;
    memset(out, 0, sizeof(*out));
    double t0 = now_s();
    SeqScanState *oscan = ExecInitSeqScan(outer, plan->n_preds, plan->preds);
    SeqScanState *iscan = NULL;
    HashJoinState *hj = NULL;
    PlanState *child = &oscan->ps;
    int has_join = plan->outer_key_col >= 0 && join && inner;
    orc_attr child_attrs[64 + 1 + GX_MAX_PAYLOAD];
    for (int i = 0; i < outer->natts; i++) child_attrs[i] = outer->attrs[i];
    if (has_join) {
        iscan = ExecInitSeqScan(inner, join->n_inner_preds, join->inner_preds);
        hj = ExecInitHashJoin(&oscan->ps, outer, plan->outer_key_col, &iscan->ps, inner, join);
        child = &hj->ps;
        child_attrs[outer->natts] = inner->attrs[join->inner_key_col];
        for (int i = 0; i < join->n_payload; i++)
            child_attrs[outer->natts + 1 + i] = inner->attrs[join->payload_cols[i]];
    }
    AggState st; memset(&st, 0, sizeof(st));
    st.child = child; st.ngroupcols = plan->n_group_cols; st.naggs = plan->n_aggs;
    st.child_attrs = child_attrs;
    st.hash_iv = 0;
    for (int c = 0; c < plan->n_group_cols; c++) {
        int att = plan->group_cols[c].side == 0 ? plan->group_cols[c].col
                                                : outer->natts + 1 + plan->group_cols[c].col;
        st.grp_att[c] = att;
        st.grp_attrs[c] = child_attrs[att];
        st.grp_attrs[c].attcacheoff = -1;
        st.grp_eq[c] = cmp_for_type(child_attrs[att].type);
        if (hj && att < outer->natts) hj->outer_needed[att] = 1;
    }
    for (int a = 0; a < plan->n_aggs; a++) {
        st.trans[a].fn = plan->aggs[a].fn;
        st.trans[a].expr = plan->aggs[a].arg;
        st.trans[a].argcol = plan->aggs[a].arg.nops == 1 ? plan->aggs[a].arg.ops[0].col : -1;
        for (int k = 0; k < plan->aggs[a].arg.nops; k++)
            if (plan->aggs[a].arg.ops[k].op == GX_OP_COL && hj)
                hj->outer_needed[plan->aggs[a].arg.ops[k].col] = 1;
    }
    sh_compute_parameters(&st.tb, 256);
    st.tb.data = (TupleHashEntryData *) calloc(st.tb.size, sizeof(TupleHashEntryData));
    st.keyslot.natts = st.ngroupcols; st.keyslot.attrs = st.grp_attrs;
    st.keyslot.values = (int64_t *) calloc(GX_MAX_GROUP_COLS, 8);
    st.keyslot.isnull = (uint8_t *) calloc(GX_MAX_GROUP_COLS, 1);

    /* agg_fill_hash_table, nodeAgg.c:2609-2648 */
    TupleHashEntryData *plain = NULL;
    if (st.ngroupcols == 0) {
        /* AGG_PLAIN: one group even over zero rows (nodeAgg.c:2320 agg_retrieve_direct) */
        int64_t k = 0; uint8_t n = 0;
        plain = LookupTupleHashEntry(&st, &k, &n);
    }
    for (;;) {
        orc_slot *slot = ExecProcNode(child);             /* fetch_input_tuple */
        if (TupIsNull(slot)) break;
        TupleHashEntryData *e = plain;
        if (!e) {
            int64_t keys[GX_MAX_GROUP_COLS]; uint8_t nulls[GX_MAX_GROUP_COLS];
            for (int c = 0; c < st.ngroupcols; c++)       /* prepare_hash_slot, nodeAgg.c:1273 */
                keys[c] = orc_slot_getattr(slot, st.grp_att[c], &nulls[c]);
            e = LookupTupleHashEntry(&st, keys, nulls);   /* lookup_hash_entries */
        }
        advance_aggregates(&st, e->additional, slot);
        if (g_err) break;
    }
    if (!g_err) fill_result(&st, out);
    g_last_secs = now_s() - t0;
    free(st.tb.data); arena_free(&st.tablecxt);
    free(st.keyslot.values); free(st.keyslot.isnull);
    if (hj) ExecEndHashJoin(hj);
    if (iscan) ExecEndSeqScan(iscan);
    ExecEndSeqScan(oscan);
    return g_err;
}

int64_t orc_exec_join(const orc_rel *outer, int outer_key_col, int n_outer_preds,
                      const gx_pred *outer_preds, const orc_rel *inner,
                      const orc_join_spec *join, int n_out_outer,
                      const int32_t *out_outer_cols, int64_t *const *cols_out)
{
    g_err = 0;
    double t0 = now_s();
    SeqScanState *oscan = ExecInitSeqScan(outer, n_outer_preds, outer_preds);
    SeqScanState *iscan = ExecInitSeqScan(inner, join->n_inner_preds, join->inner_preds);
    HashJoinState *hj = ExecInitHashJoin(&oscan->ps, outer, outer_key_col, &iscan->ps, inner, join);
    for (int c = 0; c < n_out_outer; c++) hj->outer_needed[out_outer_cols[c]] = 1;
    int64_t n = 0;
    for (;;) {
        orc_slot *s = ExecProcNode(&hj->ps);
        if (TupIsNull(s)) break;
        if (cols_out) {
            for (int c = 0; c < n_out_outer; c++)
                cols_out[c][n] = s->isnull[out_outer_cols[c]] ? INT64_MIN : s->values[out_outer_cols[c]];
            for (int p = 0; p < join->n_payload; p++) {
                int a = outer->natts + 1 + p;
                cols_out[n_out_outer + p][n] = s->isnull[a] ? INT64_MIN : s->values[a];
            }
        }
        n++;
    }
    g_last_secs = now_s() - t0;
    ExecEndHashJoin(hj); ExecEndSeqScan(iscan); ExecEndSeqScan(oscan);
    return n;
}

/* ExecChooseHashTableSize's batching decision (nodeHash.c:864-1043), reported
 * only: with hash_mem = work_mem 65535 kB * hash_mem_multiplier 1.0 the
 * reference would split into this many batches. */
void orc_last_hash_stats(int64_t *nbuckets, int64_t *ntuples, int64_t *space_used,
                         int *nbatch_if_default_work_mem)
{
    if (nbuckets) *nbuckets = g_hs_nbuckets;
    if (ntuples) *ntuples = g_hs_ntuples;
    if (space_used) *space_used = g_hs_space;
    if (nbatch_if_default_work_mem) {
        int64_t hash_mem = 65535LL * 1024;
        int nb = 1;
        while ((double) g_hs_space / nb > (double) hash_mem) nb <<= 1;
        *nbatch_if_default_work_mem = nb;
    }
}

/* Finalize HashAggregate over partial states shipped from the datanodes
 * (combine functions: int8pl, float8pl, float8_combine float.c:2725-2820) */
typedef struct comb_ent { int used; int64_t keys[GX_MAX_GROUP_COLS]; uint8_t knull[GX_MAX_GROUP_COLS];
                          double st[GX_MAX_AGGS][3]; uint8_t have[GX_MAX_AGGS]; } comb_ent;
int orc_combine_results(const gx_agg_plan *plan, const orc_result *parts, int nparts, orc_result *out)
{
    int ng = plan->n_group_cols, na = plan->n_aggs;
    int64_t total = 0;
    for (int p = 0; p < nparts; p++) total += parts[p].ngroups;
    uint64_t cap = 16; while (cap < (uint64_t) total * 2 + 2) cap <<= 1;
    comb_ent *tab = (comb_ent *) calloc(cap, sizeof(comb_ent));
    int64_t ngroups = 0;
    g_err = 0;
    for (int p = 0; p < nparts; p++) {
        const orc_result *r = &parts[p];
        for (int64_t g = 0; g < r->ngroups; g++) {
            const int64_t *k = &r->keys[g * ng];
            const uint8_t *kn = &r->nulls[g * (ng + na)];
            uint32_t h = 0;
            for (int c = 0; c < ng; c++) h = orc_hash_combine(h, kn[c] ? 0 : orc_hashint8(k[c]));
            uint64_t i = orc_murmurhash32(h) & (cap - 1);
            for (;; i = (i + 1) & (cap - 1)) {
                if (!tab[i].used) {
                    tab[i].used = 1; ngroups++;
                    for (int c = 0; c < ng; c++) { tab[i].keys[c] = k[c]; tab[i].knull[c] = kn[c]; }
                    break;
                }
                int same = 1;
                for (int c = 0; c < ng; c++) if (tab[i].knull[c] != kn[c] || (!kn[c] && tab[i].keys[c] != k[c])) same = 0;
                if (same) break;
            }
            for (int a = 0; a < na; a++) {
                const double *s = &r->states[(g * na + a) * 3];
                double *d = tab[i].st[a];
                switch (plan->aggs[a].fn) {
                    case GX_AGG_COUNT_STAR: case GX_AGG_COUNT: {      /* int8pl */
                        int64_t x, y; memcpy(&x, &d[0], 8); memcpy(&y, &s[0], 8);
                        x += y; memcpy(&d[0], &x, 8); break;
                    }
                    case GX_AGG_SUM_I4: case GX_AGG_SUM_I8: {          /* int8pl, strict */
                        if (s[1] != 0.0) break;                        /* partial is NULL */
                        int64_t x, y; memcpy(&x, &d[0], 8); memcpy(&y, &s[0], 8);
                        if (!tab[i].have[a]) { x = y; tab[i].have[a] = 1; }
                        else if (__builtin_add_overflow(x, y, &x)) g_err = GX_ERR_OVERFLOW;
                        memcpy(&d[0], &x, 8); break;
                    }
                    case GX_AGG_AVG_F8: {                              /* float8_combine */
                        double N1 = d[0], Sx1 = d[1], Sxx1 = d[2], N2 = s[0], Sx2 = s[1], Sxx2 = s[2];
                        if (N1 == 0.0) { d[0] = N2; d[1] = Sx2; d[2] = Sxx2; }
                        else if (N2 == 0.0) { }
                        else {
                            double N = N1 + N2, Sx = Sx1 + Sx2, tmp = Sx1 / N1 - Sx2 / N2;
                            double Sxx = Sxx1 + Sxx2 + N1 * N2 * tmp * tmp / N;
                            if (isinf(Sx) && !isinf(Sx1) && !isinf(Sx2)) g_err = GX_ERR_OVERFLOW;
                            d[0] = N; d[1] = Sx; d[2] = Sxx;
                        }
                        break;
                    }
                    default: {                                         /* float8pl / smaller / larger, strict */
                        if (s[0] == 0.0) break;                        /* partial is NULL */
                        if (!tab[i].have[a]) { d[1] = s[1]; tab[i].have[a] = 1; break; }
                        if (plan->aggs[a].fn == GX_AGG_SUM_F8) d[1] = float8pl(d[1], s[1]);
                        else if (plan->aggs[a].fn == GX_AGG_MIN_F8) d[1] = float8_cmp(d[1], s[1]) < 0 ? d[1] : s[1];
                        else d[1] = float8_cmp(d[1], s[1]) > 0 ? d[1] : s[1];
                    }
                }
            }
        }
    }
    memset(out, 0, sizeof(*out));
    out->n_group_cols = ng; out->n_aggs = na; out->ngroups = ngroups;
    out->keys = (int64_t *) calloc((size_t) (ngroups * (ng ? ng : 1) + 1), 8);
    out->aggs = (double *) calloc((size_t) (ngroups * (na ? na : 1) + 1), 8);
    out->nulls = (uint8_t *) calloc((size_t) (ngroups * (ng + na) + 1), 1);
    out->states = (double *) calloc((size_t) (ngroups * (na ? na : 1) * 3 + 1), 8);
    int64_t g = 0;
    for (uint64_t i = 0; i < cap; i++) {
        if (!tab[i].used) continue;
        for (int c = 0; c < ng; c++) { out->keys[g * ng + c] = tab[i].keys[c]; out->nulls[g * (ng + na) + c] = tab[i].knull[c]; }
        for (int a = 0; a < na; a++) {
            double *d = tab[i].st[a]; double res = 0.0; uint8_t nl = 0;
            switch (plan->aggs[a].fn) {
                case GX_AGG_COUNT_STAR: case GX_AGG_COUNT: res = d[0]; break;
                case GX_AGG_SUM_I4: case GX_AGG_SUM_I8: nl = !tab[i].have[a]; res = d[0]; break;
                case GX_AGG_AVG_F8: if (d[0] == 0.0) nl = 1; else res = d[1] / d[0]; break;
                default: nl = !tab[i].have[a]; res = d[1];
            }
            out->aggs[g * na + a] = nl ? 0.0 : res;
            out->nulls[g * (ng + na) + ng + a] = nl;
            memcpy(&out->states[(g * na + a) * 3], d, 24);
        }
        g++;
    }
    free(tab);
    return g_err;
}
