/*
 * otb_oracle.h — CPU restatement of the reference's executor hot path.
 *
 * *** TEST INFRASTRUCTURE.  NOT PART OF THE PRODUCT. ***
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load this library.  libgpuexec.so never links,
 * loads or calls it; the GPU path has no CPU fallback.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Parity status: the hash functions, tuple
 * forming/deforming, page layout, bloom filter, routing hash and float8
 * transition functions are PINNED against the reference's own object code
 * (oracle/ref/Makefile compiles the reference's leaf .c files from where they
 * lie into oracle/_ref/libotbref.so; tests/test_oracle_vs_ref.py) and against
 * the reference's SQL goldens (tests/golden/).  The executor nodes themselves
 * (nodeHash.c/nodeHashjoin.c/nodeAgg.c) cannot be linked without the whole
 * backend; for them the restatement is argued from file:line and pinned only
 * by the SQL-level goldens.
 */
#ifndef OTB_ORACLE_H
#define OTB_ORACLE_H

#include <stdint.h>
#include <stddef.h>
#include "../include/gpuexec.h"      /* plan descriptor PODs only */

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ hashes */
uint32_t orc_hash_any(const unsigned char *k, int keylen);      /* hashfunc.c:619 */
uint32_t orc_hash_uint32(uint32_t k);                           /* hashfunc.c:1044 */
uint32_t orc_hashint4(int32_t v);                               /* hashfunc.c:80  */
uint32_t orc_hashint8(int64_t v);                               /* hashfunc.c:92  */
uint32_t orc_hashchar(int8_t v);                                /* hashfunc.c:48  */
uint32_t orc_hashfloat8(double v);                              /* hashfunc.c:325 */
uint32_t orc_crc32c(uint32_t crc, const void *data, size_t len);/* port/pg_crc32c_sb8.c */
uint32_t orc_hash_any_new(const unsigned char *k, int keylen);  /* hashfunc.c:112 */
uint32_t orc_hashint4new(int32_t v);                            /* hashfunc.c:150 */
uint32_t orc_hashint8new(int64_t v);                            /* hashfunc.c:168 */
uint32_t orc_hashcharnew(int8_t v);                             /* hashfunc.c:136 */
uint32_t orc_hashfloat8new(double v);                           /* hashfunc.c:231 */
uint32_t orc_murmurhash32(uint32_t h);                          /* hashutils.h:39 */
uint32_t orc_hash_combine(uint32_t a, uint32_t b);              /* hashutils.h:16 */

/* type-dispatched per-column hash (type = GX_INT4...) */
uint32_t orc_hash_datum(int type, int64_t datum);               /* Jenkins family */
uint32_t orc_hash_datum_new(int type, int64_t datum);           /* CRC32C family  */

/* ----------------------------------------------------------------- routing */
/* EvaluateHashkey, pgxc/locator/locator.c:1611-1628 */
uint32_t orc_evaluate_hashkey(const int *types, const uint8_t *isnull,
                              const int64_t *datums, int natts);
/* GetNodeIndexByHashValue, pgxc/shard/shardmap.c:1147-1160 */
int32_t  orc_shard_index(uint32_t hashvalue);
/* default map, catalog/pgxc_shard_map.c:93 */
void     orc_default_shardmap(int32_t *map /* 4096 */, int nnodes);
int32_t  orc_route_node(const int32_t *shardmap, int type, int64_t datum, int isnull);

/* --------------------------------------------------------------- generator */
/* columns (may be NULL to skip) receive rows for orders [order0, order1);
 * node filter as in gx_table_generate.  Returns rows written. */
int64_t orc_gen_orders(int sf, int64_t order0, int64_t order1, int node, int nnodes,
                       int64_t *o_orderkey, int32_t *o_custkey, int32_t *o_orderdate,
                       int32_t *o_shippriority);
int64_t orc_gen_lineitem_count(int sf, int64_t order0, int64_t order1, int node, int nnodes);
int64_t orc_gen_lineitem(int sf, int64_t order0, int64_t order1, int node, int nnodes,
                         int64_t *l_orderkey, double *l_quantity, double *l_extendedprice,
                         double *l_discount, double *l_tax, int32_t *l_shipdate,
                         int8_t *l_returnflag, int8_t *l_linestatus);
int64_t orc_gen_customer(int sf, int64_t c0, int64_t c1, int node, int nnodes,
                         int32_t *c_custkey, int8_t *c_mktsegment);

/* ------------------------------------------------------- heap relation */
typedef struct orc_rel orc_rel;
/* A relation of by-value attributes stored as OpenTenBase heap pages
 * (bufpage.h:153-175, htup_details.h:126-201, itemid.h). */
orc_rel *orc_rel_create(int natts, const int32_t *types);
int      orc_rel_add_column(orc_rel *r, int32_t type);   /* ALTER TABLE ADD COLUMN, no rewrite */
void     orc_rel_free(orc_rel *r);
/* heap_form_tuple + PageAddItem (heaptuple.c:1012, bufpage.c:218) */
int      orc_rel_insert_columns(orc_rel *r, const void *const *cols,
                                const uint8_t *const *nulls, int64_t nrows);
int64_t  orc_rel_ntuples(const orc_rel *r);
int64_t  orc_rel_npages(const orc_rel *r);
const void *orc_rel_page(const orc_rel *r, int64_t pageno);  /* 8192 bytes */
/* copies all pages contiguously into out (npages*8192 bytes) */
void     orc_rel_copy_pages(const orc_rel *r, void *out);
/* mark a tuple deleted (xmax committed) so the visibility stub skips it */
int      orc_rel_delete_tuple(orc_rel *r, int64_t pageno, int lineoff /* 1-based */);
/* heapgetpage + slot_deform_tuple: scan into columns; returns visible rows */
int64_t  orc_rel_scan_columns(const orc_rel *r, int ncols, const int32_t *attnums,
                              void *const *cols_out, uint8_t *const *nulls_out);

/* ------------------------------------------------------- forward-node pages */
/* The redistribute wire format (forward/fnbufpage.h:54-65): 8192-byte pages of MAXALIGNed minimal tuples.
 * pack follows the sender (FragmentSendAttrs + FragmentGetPage's fit rule, execFragment.c:2067-2136, :1857-1876;
 * end_marker adds FragmentSendNullTuple's MAX_UINT32 word + FNPAGE_END); returns the number of pages,
 * -1 if cap_pages is too small, -2 for a tuple that would need FragmentSendHuge.  unpack follows the receiver's
 * iterator (fnbufpage.h:107-127, tqueueThread.c:913-925) and slot_deform_tuple; returns the rows, -1 / -2 likewise. */
typedef struct orc_fnpage_id { int64_t qid_timestamp_nodeid, qid_sequence; uint16_t fid, nodeid, workerid; uint8_t virtualid, pad; } orc_fnpage_id;
int64_t orc_fnpage_pack(int natts, const int32_t *types, const void *const *cols, const uint8_t *const *nulls, int64_t nrows,
                        const orc_fnpage_id *id, int end_marker, uint8_t *pages, int64_t cap_pages);
int64_t orc_fnpage_unpack(const uint8_t *pages, int64_t npages, int natts, const int32_t *types,
                          void *const *cols_out, uint8_t *const *nulls_out, int64_t cap_rows);

/* ------------------------------------------------------- executor */
typedef struct orc_result {
    int32_t  n_group_cols, n_aggs;
    int64_t  ngroups;
    int64_t *keys;      /* ngroups * n_group_cols (group col widened to int64) */
    double  *aggs;      /* ngroups * n_aggs (int64 results bit-cast)           */
    uint8_t *nulls;     /* ngroups * (n_group_cols + n_aggs)                   */
    /* partial (transition) states for combine tests:
     * per group per agg: {N or count (bitcast int64 for count aggs), Sx, Sxx} */
    double  *states;    /* ngroups * n_aggs * 3 */
} orc_result;
void orc_result_free(orc_result *r);

typedef struct orc_join_spec {
    int32_t  inner_key_col;
    int32_t  n_inner_preds;
    gx_pred  inner_preds[GX_MAX_PREDS];
    int32_t  n_payload;
    int32_t  payload_cols[GX_MAX_PAYLOAD];
    int32_t  inner_unique;
    int32_t  jointype;          /* JoinType (nodes/nodes.h): 0 INNER, 1 LEFT, 4 SEMI, 5 ANTI */
} orc_join_spec;

/* Volcano, one tuple per ExecProcNode call:
 * SeqScan(outer)[quals] [-> HashJoin(Hash(SeqScan(inner)[quals]))] -> Agg(HASHED)
 * (nodeSeqscan.c:60, execScan.c:141, nodeHash.c:157, nodeHashjoin.c:186,
 *  nodeAgg.c:2212).  join may be NULL.  node_id/nnodes feed the partial-agg
 * hash_iv (execGrouping.c:207-210); pass -1 for a plain aggregate. */
int orc_exec_agg(const orc_rel *outer, const orc_rel *inner, const orc_join_spec *join,
                 const gx_agg_plan *plan, orc_result *out);

/* HashJoin materialised: out columns = out_outer_cols ++ payload cols;
 * returns rows; cols_out[c] must hold outer-ntuples*max-fanout rows — call
 * with cols_out == NULL first to get the count. */
int64_t orc_exec_join(const orc_rel *outer, int outer_key_col, int n_outer_preds,
                      const gx_pred *outer_preds, const orc_rel *inner,
                      const orc_join_spec *join, int n_out_outer,
                      const int32_t *out_outer_cols, int64_t *const *cols_out);

/* combine partial results from several datanodes (Finalize HashAggregate:
 * int8pl / float8pl / float8_combine, float.c:2725) and finalize. */
int orc_combine_results(const gx_agg_plan *plan, const orc_result *parts, int nparts,
                        orc_result *out);

/* timing hook for bench.py's cpu_baseline: wall seconds of the last
 * orc_exec_agg / orc_exec_join call, excluding relation construction */
double orc_last_exec_seconds(void);

/* hash-join instrumentation of the last run (ExecHashAccumInstrumentation) */
void orc_last_hash_stats(int64_t *nbuckets, int64_t *ntuples, int64_t *space_used,
                         int *nbatch_if_default_work_mem);

/* ------------------------------------------------------- bloom filter */
typedef struct orc_bloom orc_bloom;
orc_bloom *orc_bloom_create(int64_t ndv_estimate);   /* bloomfilter.c:54  */
void       orc_bloom_insert(orc_bloom *b, uint32_t hash); /* :140 */
int        orc_bloom_find(const orc_bloom *b, uint32_t hash); /* :162 */
int        orc_bloom_log_num_buckets(const orc_bloom *b);
const uint32_t *orc_bloom_words(const orc_bloom *b, int64_t *nwords);
void       orc_bloom_free(orc_bloom *b);

#ifdef __cplusplus
}
#endif
#endif
