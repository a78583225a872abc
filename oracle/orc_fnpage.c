/*
 * orc_fnpage.c — forward-node pages: the wire format of a redistribute.
 * TEST INFRASTRUCTURE (see otb_oracle.h).
 *
 * Restates
 *   FnPageHeaderData / iterator macros      src/include/forward/fnbufpage.h:54-65, :107-127
 *   FnPageInit                              src/backend/forward/storage/fnbufpage.c:30-42
 *   FragmentSendAttrs                       src/backend/executor/execFragment.c:2067-2136
 *   FragmentGetPage's "does it fit" rule    src/backend/executor/execFragment.c:1857-1876
 *   FragmentSendNullTuple (end of stream)   src/backend/executor/execFragment.c:1963-1975
 *   heap_minimal_tuple_header_size / heap_form_minimal_tuple_ptr
 *                                           src/backend/access/common/heaptuple.c:1819-1895
 *   MinimalTupleData                        src/include/access/htup_details.h:744-784
 *   the receiver's walk                     src/backend/executor/tqueueThread.c:913-925
 * Pinned against the reference's own object code (heaptuple.o, fnbufpage.o) by tests/test_oracle_vs_ref.py.
 *
 * A page: 32-byte header {lower u32, fid u16, nodeid u16, queryid 2 x i64, flag u32, workerid u16, virtualid u8, pad u8},
 * then MAXALIGNed minimal tuples back to back up to `lower`.  A minimal tuple is a heap tuple without its first 32
 * bytes: {t_len u32, 2 pad bytes, t_infomask2 u16, t_infomask u16, t_infomask3 u16, t_shardid u16, t_hoff u8, t_bits[]};
 * t_hoff still counts the 32 missing bytes.
 */
#include <stdlib.h>
#include <string.h>
#include "orc_internal.h"

#define FNP_HDR          32
#define FNP_LOWER        0
#define FNP_FID          4
#define FNP_NODEID       6
#define FNP_QID_TS       8
#define FNP_QID_SEQ      16
#define FNP_FLAG         24
#define FNP_WORKERID     28
#define FNP_VIRTUALID    30
#define FNPAGE_HUGE      1u
#define FNPAGE_END       8u
#define MT_HDR           15            /* SizeofMinimalTupleHeader */
#define MT_INFOMASK2     6
#define MT_INFOMASK      8
#define MT_SHARDID       12
#define MT_HOFF          14
#define MT_BITS          15
#define INVALID_SHARDID  4096          /* postgres_ext.h:80-81 */

static inline void wr16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }
static inline void wr32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
static inline void wr64(uint8_t *p, int64_t v) { memcpy(p, &v, 8); }
static inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

static int64_t col_value(int type, const void *col, int64_t i)
{
    switch (type) {
        case GX_INT4: case GX_DATE:   return (int64_t) ((const int32_t *) col)[i];
        case GX_INT8: case GX_FLOAT8: return ((const int64_t *) col)[i];
        default:                      return (int64_t) ((const int8_t *) col)[i];
    }
}
static void store_col(int type, void *col, int64_t i, int64_t v)
{
    switch (type) {
        case GX_INT4: case GX_DATE:   ((int32_t *) col)[i] = (int32_t) v; break;
        case GX_INT8: case GX_FLOAT8: ((int64_t *) col)[i] = v; break;
        default:                      ((int8_t *) col)[i] = (int8_t) v; break;
    }
}

static uint8_t *page_start(uint8_t *pages, int64_t *npages, int64_t cap, const orc_fnpage_id *id)
{
    if (*npages >= cap) return NULL;
    uint8_t *pg = pages + (*npages)++ * ORC_BLCKSZ;
    memset(pg, 0, ORC_BLCKSZ);
    wr32(pg + FNP_LOWER, FNP_HDR);                          /* FnPageInit */
    wr16(pg + FNP_FID, id->fid); wr16(pg + FNP_NODEID, id->nodeid);
    wr64(pg + FNP_QID_TS, id->qid_timestamp_nodeid); wr64(pg + FNP_QID_SEQ, id->qid_sequence);
    wr32(pg + FNP_FLAG, 0); wr16(pg + FNP_WORKERID, id->workerid);
    pg[FNP_VIRTUALID] = id->virtualid;                      /* FragmentGetPage, execFragment.c:1838 */
    return pg;
}

int64_t orc_fnpage_pack(int natts, const int32_t *types, const void *const *cols, const uint8_t *const *nulls, int64_t nrows,
                        const orc_fnpage_id *id, int end_marker, uint8_t *pages, int64_t cap_pages)
{
    orc_attr attrs[64];
    int64_t values[64]; uint8_t isnull[64];
    if (natts > 64) return -1;
    for (int i = 0; i < natts; i++) { attrs[i].type = types[i]; orc_type_layout(types[i], &attrs[i].attlen, &attrs[i].attalign); attrs[i].attcacheoff = -1; }
    int64_t npages = 0;
    uint8_t *pg = NULL;
    for (int64_t r = 0; r < nrows; r++) {
        int hasnull = 0, hasvar = 0;
        for (int i = 0; i < natts; i++) {
            isnull[i] = nulls && nulls[i] ? nulls[i][r] != 0 : 0;
            values[i] = isnull[i] ? 0 : col_value(types[i], cols[i], r);
            hasnull |= isnull[i];
            if (attrs[i].attlen == -1 && !isnull[i]) hasvar = 1;
        }
        /* heap_minimal_tuple_header_size: MAXALIGN(header + bitmap) */
        const uint32_t hoff = (uint32_t) ORC_MAXALIGN(MT_HDR + (hasnull ? (natts + 7) / 8 : 0));
        const uint32_t len = hoff + orc_compute_data_size(attrs, natts, values, isnull);
        const uint32_t alen = (uint32_t) ORC_MAXALIGN(len);
        if (alen > ORC_BLCKSZ - FNP_HDR) return -2;                               /* FragmentSendHuge: not restated */
        if (!pg || alen > ORC_BLCKSZ - rd32(pg + FNP_LOWER)) {                   /* FragmentGetPage: full -> next page */
            pg = page_start(pages, &npages, cap_pages, id);
            if (!pg) return -1;
        }
        uint8_t *t = pg + rd32(pg + FNP_LOWER);
        memset(t, 0, len);                                                       /* heap_form_minimal_tuple_ptr */
        wr32(t, len);
        wr16(t + MT_INFOMASK2, (uint16_t) (natts & HEAP_NATTS_MASK));
        t[MT_HOFF] = (uint8_t) (hoff + ORC_MINIMAL_TUPLE_OFFSET);
        wr16(t + MT_INFOMASK, (uint16_t) ((hasnull ? HEAP_HASNULL : 0) | (hasvar ? HEAP_HASVARWIDTH : 0)));   /* heap_fill_tuple */
        orc_form_data(attrs, natts, values, isnull, hasnull ? t + MT_BITS : NULL, t + hoff);
        wr16(t + MT_SHARDID, INVALID_SHARDID);
        wr32(pg + FNP_LOWER, rd32(pg + FNP_LOWER) + alen);
    }
    if (end_marker) {                                                            /* FragmentSendNullTuple */
        if (!pg || 4 > ORC_BLCKSZ - rd32(pg + FNP_LOWER)) {
            pg = page_start(pages, &npages, cap_pages, id);
            if (!pg) return -1;
        }
        wr32(pg + rd32(pg + FNP_LOWER), 0xFFFFFFFFu);
        wr32(pg + FNP_LOWER, rd32(pg + FNP_LOWER) + 4);
        wr32(pg + FNP_FLAG, rd32(pg + FNP_FLAG) | FNPAGE_END);
    }
    return npages;
}

int64_t orc_fnpage_unpack(const uint8_t *pages, int64_t npages, int natts, const int32_t *types,
                          void *const *cols_out, uint8_t *const *nulls_out, int64_t cap_rows)
{
    orc_attr attrs[64];
    int64_t values[64]; uint8_t isnull[64];
    if (natts > 64) return -1;
    for (int i = 0; i < natts; i++) { attrs[i].type = types[i]; orc_type_layout(types[i], &attrs[i].attlen, &attrs[i].attalign); attrs[i].attcacheoff = -1; }
    int64_t rows = 0;
    for (int64_t p = 0; p < npages; p++) {
        const uint8_t *pg = pages + p * ORC_BLCKSZ;
        if (rd32(pg + FNP_FLAG) & FNPAGE_HUGE) return -2;
        const uint32_t lower = rd32(pg + FNP_LOWER);
        uint32_t off = FNP_HDR;                                                  /* InitFnPageIterator */
        while (off < lower) {                                                    /* FnPageIterateDone */
            const uint32_t len = rd32(pg + off);                                 /* FnPageIterateNext */
            if (len == 0xFFFFFFFFu) break;
            if (rows >= cap_rows) return -1;
            orc_slot slot; memset(&slot, 0, sizeof(slot));
            slot.natts = natts; slot.attrs = attrs; slot.values = values; slot.isnull = isnull;
            slot.tuple = pg + off - ORC_MINIMAL_TUPLE_OFFSET;                    /* ExecStoreMinimalTuple's view of it */
            orc_slot_deform(&slot, natts);
            for (int i = 0; i < natts; i++) {
                if (nulls_out && nulls_out[i]) nulls_out[i][rows] = isnull[i];
                store_col(types[i], cols_out[i], rows, isnull[i] ? 0 : values[i]);
            }
            rows++;
            off += (uint32_t) ORC_MAXALIGN(len);
        }
    }
    return rows;
}
