/*
 * orc_heap.c — OpenTenBase heap pages: form, store, scan, deform.
 * TEST INFRASTRUCTURE (see otb_oracle.h).
 *
 * Restates, for by-value attributes and short varlenas:
 *   PageInit / PageAddItemExtended     src/backend/storage/page/bufpage.c:42,218
 *   heap_compute_data_size / heap_fill_tuple / heap_form_tuple
 *                                      src/backend/access/common/heaptuple.c:104,255,1012
 *   heapgetpage (visibility pass)      src/backend/access/heap/heapam.c:388-513
 *   slot_deform_tuple                  src/backend/access/common/heaptuple.c:1518-1614
 *   att_align_nominal/att_align_pointer/fetchatt/att_addlength_pointer
 *                                      src/include/access/tupmacs.h
 */
#include <stdlib.h>
#include <string.h>
#include "orc_internal.h"

void orc_type_layout(int type, int16_t *attlen, int8_t *attalign)
{
    switch (type) {
        case GX_INT4:
        case GX_DATE:    *attlen = 4; *attalign = 4; break;   /* typalign 'i' */
        case GX_INT8:
        case GX_FLOAT8:  *attlen = 8; *attalign = 8; break;   /* typalign 'd' */
        case GX_CHAR:    *attlen = 1; *attalign = 1; break;   /* typalign 'c' */
        case ORC_BPCHAR1:*attlen = -1; *attalign = 4; break;  /* varlena, 'i' */
        default: abort();
    }
}

orc_rel *orc_rel_create(int natts, const int32_t *types)
{
    orc_rel *r = (orc_rel *) calloc(1, sizeof(*r));
    r->natts = natts;
    r->attrs = (orc_attr *) calloc((size_t) natts, sizeof(orc_attr));
    for (int i = 0; i < natts; i++) {
        r->attrs[i].type = types[i];
        orc_type_layout(types[i], &r->attrs[i].attlen, &r->attrs[i].attalign);
        r->attrs[i].attcacheoff = -1;
    }
    return r;
}

/* ALTER TABLE ADD COLUMN without a rewrite: the descriptor grows, tuples already on the
 * pages keep their old attribute count in t_infomask2 and read the new attribute as NULL
 * (no missing-value default is modelled: heap_deform_tuple, heaptuple.c:1424,1497-1502 with
 * atthasmissing = false). */
int orc_rel_add_column(orc_rel *r, int32_t type)
{
    if (r->natts >= 64) return -1;
    r->attrs = (orc_attr *) realloc(r->attrs, (size_t) (r->natts + 1) * sizeof(orc_attr));
    memset(&r->attrs[r->natts], 0, sizeof(orc_attr));
    r->attrs[r->natts].type = type;
    orc_type_layout(type, &r->attrs[r->natts].attlen, &r->attrs[r->natts].attalign);
    r->attrs[r->natts].attcacheoff = -1;
    r->natts++;
    return 0;
}

void orc_rel_free(orc_rel *r)
{
    if (!r) return;
    for (int64_t i = 0; i < r->npages; i++) free(r->pages[i]);
    free(r->pages);
    free(r->attrs);
    free(r);
}

int64_t orc_rel_ntuples(const orc_rel *r) { return r->ntuples; }
int64_t orc_rel_npages(const orc_rel *r) { return r->npages; }
const void *orc_rel_page(const orc_rel *r, int64_t p) { return r->pages[p]; }
void orc_rel_copy_pages(const orc_rel *r, void *out)
{
    for (int64_t i = 0; i < r->npages; i++)
        memcpy((uint8_t *) out + i * ORC_BLCKSZ, r->pages[i], ORC_BLCKSZ);
}

static inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline void wr16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline void wr32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }

/* PageHeaderData field offsets (bufpage.h:153-175 with _SHARDING_, _MLS_, GTS and
 * __OPENTENBASE_C__ on: LocationIndex is uint32, bufpage.h:85-89).  Verified
 * against the reference's own headers by tests/test_oracle_vs_ref.py:
 * pd_lsn 0, pd_checksum 8, pd_flags 10, pd_shard 12, pd_lower 16, pd_upper 20,
 * pd_special 24, pd_pagesize_version 28, pd_algorithm_id 30, pd_prune_ts 32,
 * pd_prune_xid 40, pd_linp 44. */
#define PD_FLAGS    10
#define PD_SHARD    12
#define PD_LOWER    16
#define PD_UPPER    20
#define PD_SPECIAL  24
#define PD_PAGESIZE_VERSION 28

/* PageInit, bufpage.c:42-60 */
static uint8_t *page_new(orc_rel *r)
{
    uint8_t *pg;
    if (posix_memalign((void **) &pg, 64, ORC_BLCKSZ)) abort();
    memset(pg, 0, ORC_BLCKSZ);
    wr32(pg + PD_LOWER, ORC_PAGE_HDR);
    wr32(pg + PD_UPPER, ORC_BLCKSZ);
    wr32(pg + PD_SPECIAL, ORC_BLCKSZ);
    wr16(pg + PD_PAGESIZE_VERSION, ORC_BLCKSZ | 4);   /* PG_PAGE_LAYOUT_VERSION 4 */
    if (r->npages == r->pages_cap) {
        r->pages_cap = r->pages_cap ? r->pages_cap * 2 : 64;
        r->pages = (uint8_t **) realloc(r->pages, (size_t) r->pages_cap * sizeof(uint8_t *));
    }
    r->pages[r->npages++] = pg;
    return pg;
}

/* att_align_nominal, tupmacs.h */
static inline uint32_t align_nominal(uint32_t off, int8_t attalign)
{
    return (off + (uint32_t) attalign - 1) & ~((uint32_t) attalign - 1);
}

/* heap_compute_data_size, heaptuple.c:104-160.  A bpchar(1) datum is a
 * 4-byte-header varlena that VARATT_CAN_MAKE_SHORT: stored with a 1-byte
 * header and no alignment padding. */
uint32_t orc_compute_data_size(const orc_attr *attrs, int natts, const int64_t *values,
                               const uint8_t *isnull)
{
    uint32_t len = 0;
    (void) values;
    for (int i = 0; i < natts; i++) {
        if (isnull && isnull[i]) continue;
        if (attrs[i].attlen == -1) {
            len += 2;                              /* 1-byte header + 1 data byte */
        } else {
            len = align_nominal(len, attrs[i].attalign);
            len += (uint32_t) attrs[i].attlen;
        }
    }
    return len;
}

/* heap_fill_tuple, heaptuple.c:255-420: null bitmap bit SET means NOT NULL */
uint32_t orc_form_data(const orc_attr *attrs, int natts, const int64_t *values,
                       const uint8_t *isnull, uint8_t *bits, uint8_t *dst)
{
    uint32_t off = 0;
    if (bits) memset(bits, 0, (size_t) ((natts + 7) / 8));
    for (int i = 0; i < natts; i++) {
        if (isnull && isnull[i]) continue;
        if (bits) bits[i >> 3] |= (uint8_t) (1 << (i & 7));
        if (attrs[i].attlen == -1) {
            dst[off++] = (uint8_t) ((2 << 1) | 0x01);   /* SET_VARSIZE_1B(len=2), postgres.h */
            dst[off++] = (uint8_t) values[i];
            continue;
        }
        uint32_t aligned = align_nominal(off, attrs[i].attalign);
        while (off < aligned) dst[off++] = 0;      /* pad bytes are zero */
        switch (attrs[i].attlen) {                 /* store_att_byval, tupmacs.h */
            case 1: dst[off] = (uint8_t) values[i]; break;
            case 2: { int16_t v = (int16_t) values[i]; memcpy(dst + off, &v, 2); break; }
            case 4: { int32_t v = (int32_t) values[i]; memcpy(dst + off, &v, 4); break; }
            case 8: memcpy(dst + off, &values[i], 8); break;
        }
        off += (uint32_t) attrs[i].attlen;
    }
    return off;
}

/* heap_form_tuple (heaptuple.c:1012-1110) straight into a page via
 * PageAddItemExtended (bufpage.c:218-380) under heap_insert's fill rule
 * (RelationGetBufferForTuple, hio.c: a page takes the tuple while
 * MAXALIGN(len) <= PageGetHeapFreeSpace). */
static void rel_insert_tuple(orc_rel *r, const int64_t *values, const uint8_t *isnull)
{
    int hasnull = 0;
    for (int i = 0; i < r->natts; i++) if (isnull[i]) hasnull = 1;
    uint32_t hoff = ORC_HEAP_HDR + (hasnull ? (uint32_t) ((r->natts + 7) / 8) : 0);
    hoff = (uint32_t) ORC_MAXALIGN(hoff);
    uint32_t dlen = orc_compute_data_size(r->attrs, r->natts, values, isnull);
    uint32_t len = hoff + dlen;
    uint32_t alen = (uint32_t) ORC_MAXALIGN(len);

    uint8_t *pg = r->npages ? r->pages[r->npages - 1] : NULL;
    if (pg) {
        uint32_t lower = rd32(pg + PD_LOWER), upper = rd32(pg + PD_UPPER);
        /* PageGetFreeSpace: space minus one new line pointer */
        int32_t freesp = (int32_t) upper - (int32_t) lower - 4;
        if (freesp < (int32_t) alen) pg = NULL;
    }
    if (!pg) pg = page_new(r);

    uint32_t lower = rd32(pg + PD_LOWER), upper = rd32(pg + PD_UPPER);
    uint32_t offnum = (lower - ORC_PAGE_HDR) / 4 + 1;           /* 1-based OffsetNumber */
    upper -= alen;
    uint8_t *tup = pg + upper;
    memset(tup, 0, alen);
    wr32(tup + HTH_XMIN, 3);                                    /* FirstNormalTransactionId */
    wr32(tup + HTH_XMAX, 0);
    /* t_ctid = (block, offset): ItemPointerData = BlockIdData{bi_hi,bi_lo} + posid */
    uint32_t blk = (uint32_t) (r->npages - 1);
    wr16(tup + HTH_CTID, (uint16_t) (blk >> 16));
    wr16(tup + HTH_CTID + 2, (uint16_t) (blk & 0xFFFF));
    wr16(tup + HTH_CTID + 4, (uint16_t) offnum);
    wr16(tup + HTH_INFOMASK2, (uint16_t) (r->natts & HEAP_NATTS_MASK));
    uint16_t infomask = HEAP_XMIN_COMMITTED | HEAP_XMAX_INVALID;
    if (hasnull) infomask |= HEAP_HASNULL;
    for (int i = 0; i < r->natts; i++) if (r->attrs[i].attlen == -1 && !isnull[i]) infomask |= HEAP_HASVARWIDTH;
    wr16(tup + HTH_INFOMASK, infomask);
    {   /* t_shardid: shard of the distribution column (column 0) */
        uint8_t n0 = isnull[0];
        int t0 = r->attrs[0].type == ORC_BPCHAR1 ? GX_CHAR : r->attrs[0].type;
        uint32_t h = orc_evaluate_hashkey(&t0, &n0, &values[0], 1);
        wr16(tup + HTH_SHARDID, (uint16_t) orc_shard_index(h));
    }
    tup[HTH_HOFF] = (uint8_t) hoff;
    orc_form_data(r->attrs, r->natts, values, isnull, hasnull ? tup + HTH_BITS : NULL, tup + hoff);

    /* ItemIdData: lp_off:15, lp_flags:2, lp_len:15 (itemid.h:24-29) */
    uint32_t lp = (upper & 0x7FFF) | ((uint32_t) LP_NORMAL << 15) | ((len & 0x7FFF) << 17);
    wr32(pg + lower, lp);
    wr32(pg + PD_LOWER, lower + 4);
    wr32(pg + PD_UPPER, upper);
    r->ntuples++;
}

static int64_t col_value(int type, const void *col, int64_t i)
{
    switch (type) {
        case GX_INT4:
        case GX_DATE:   return (int64_t) ((const int32_t *) col)[i];
        case GX_INT8:
        case GX_FLOAT8: return ((const int64_t *) col)[i];     /* float8 Datum = bit pattern */
        case GX_CHAR:
        case ORC_BPCHAR1: return (int64_t) ((const int8_t *) col)[i];
    }
    abort();
}

int orc_rel_insert_columns(orc_rel *r, const void *const *cols,
                           const uint8_t *const *nulls, int64_t nrows)
{
    int64_t values[64];
    uint8_t isnull[64];
    if (r->natts > 64) return -1;
    for (int64_t i = 0; i < nrows; i++) {
        for (int c = 0; c < r->natts; c++) {
            isnull[c] = (nulls && nulls[c]) ? nulls[c][i] : 0;
            values[c] = isnull[c] ? 0 : col_value(r->attrs[c].type, cols[c], i);
        }
        rel_insert_tuple(r, values, isnull);
    }
    return 0;
}

int orc_rel_delete_tuple(orc_rel *r, int64_t pageno, int lineoff)
{
    if (pageno < 0 || pageno >= r->npages) return -1;
    uint8_t *pg = r->pages[pageno];
    uint32_t nlines = (rd32(pg + PD_LOWER) - ORC_PAGE_HDR) / 4;
    if (lineoff < 1 || (uint32_t) lineoff > nlines) return -1;
    uint32_t lp = rd32(pg + ORC_PAGE_HDR + 4 * (uint32_t) (lineoff - 1));
    uint8_t *tup = pg + (lp & 0x7FFF);
    uint16_t im = rd16(tup + HTH_INFOMASK);
    if (!(im & HEAP_XMAX_INVALID)) return 0;           /* already deleted */
    im = (uint16_t) ((im & ~HEAP_XMAX_INVALID) | HEAP_XMAX_COMMITTED);
    wr16(tup + HTH_INFOMASK, im);
    wr32(tup + HTH_XMAX, 4);
    r->ntuples--;
    return 0;
}

/* Stand-in for HeapTupleSatisfiesMVCC (utils/time/tqual.c:1203): hint bits only. */
static inline int tuple_visible(const uint8_t *tup)
{
    uint16_t im = rd16(tup + HTH_INFOMASK);
    if (!(im & HEAP_XMIN_COMMITTED)) return 0;
    if (im & HEAP_XMAX_INVALID) return 1;
    return !(im & HEAP_XMAX_COMMITTED);
}

/* heapgetpage, heapam.c:388-513: one visibility pass per page filling
 * rs_vistuples[]; returns rs_ntuples. */
int orc_heapgetpage(const uint8_t *pg, uint16_t *vistuples)
{
    int ntup = 0;
    uint32_t lines = (rd32(pg + PD_LOWER) - ORC_PAGE_HDR) / 4;   /* PageGetMaxOffsetNumber */
    for (uint32_t lineoff = 1; lineoff <= lines; lineoff++) {
        uint32_t lp = rd32(pg + ORC_PAGE_HDR + 4 * (lineoff - 1));
        if (((lp >> 15) & 3) != LP_NORMAL) continue;             /* ItemIdIsNormal */
        const uint8_t *tup = pg + (lp & 0x7FFF);
        if (tuple_visible(tup))
            vistuples[ntup++] = (uint16_t) lineoff;
    }
    return ntup;
}

/* slot_deform_tuple, heaptuple.c:1518-1614 */
void orc_slot_deform(orc_slot *slot, int natts)
{
    const uint8_t *tup = slot->tuple;
    uint16_t infomask = rd16(tup + HTH_INFOMASK);
    int hasnulls = (infomask & HEAP_HASNULL) != 0;
    const uint8_t *bp = tup + HTH_BITS;
    const uint8_t *tp = tup + tup[HTH_HOFF];
    int attnum = slot->nvalid;
    uint32_t off;
    int slow;

    if (attnum == 0) { off = 0; slow = 0; }
    else { off = slot->off; slow = slot->slow; }

    /* natts = Min(HeapTupleHeaderGetNatts(tup), natts) (heaptuple.c:1424, :1555): attributes the
     * tuple was written without read as NULL (getmissingattr without a missing value) */
    const int want = natts;
    const int tnatts = (int) (rd16(tup + HTH_INFOMASK2) & HEAP_NATTS_MASK);
    if (natts > tnatts) natts = tnatts;
    for (int a = natts > attnum ? natts : attnum; a < want; a++) { slot->values[a] = 0; slot->isnull[a] = 1; }

    for (; attnum < natts; attnum++) {
        orc_attr *att = &slot->attrs[attnum];
        if (hasnulls && !(bp[attnum >> 3] & (1 << (attnum & 7)))) {   /* att_isnull */
            slot->values[attnum] = 0;
            slot->isnull[attnum] = 1;
            slow = 1;
            continue;
        }
        slot->isnull[attnum] = 0;
        if (!slow && att->attcacheoff >= 0)
            off = (uint32_t) att->attcacheoff;
        else if (att->attlen == -1) {
            if (!slow && off == align_nominal(off, att->attalign))
                att->attcacheoff = (int32_t) off;
            else {
                /* att_align_pointer: a non-zero byte is a 1-byte varlena header */
                if (tp[off] == 0) off = align_nominal(off, att->attalign);
                slow = 1;
            }
        } else {
            off = align_nominal(off, att->attalign);
            if (!slow) att->attcacheoff = (int32_t) off;
        }
        /* fetchatt + att_addlength_pointer */
        switch (att->attlen) {
            case 1: slot->values[attnum] = (int64_t) (int8_t) tp[off]; off += 1; break;
            case 2: { int16_t v; memcpy(&v, tp + off, 2); slot->values[attnum] = v; off += 2; break; }
            case 4: { int32_t v; memcpy(&v, tp + off, 4); slot->values[attnum] = v; off += 4; break; }
            case 8: { int64_t v; memcpy(&v, tp + off, 8); slot->values[attnum] = v; off += 8; break; }
            default: {
                /* varlena: VARSIZE_ANY (postgres.h).  bpchar(1) datum -> its one data byte */
                uint8_t h = tp[off];
                uint32_t vsz, hdr;
                if (h & 0x01) { vsz = (h >> 1) & 0x7F; hdr = 1; }
                else { vsz = (rd32(tp + off) >> 2) & 0x3FFFFFFF; hdr = 4; }
                slot->values[attnum] = (vsz > hdr) ? (int64_t) (int8_t) tp[off + hdr] : 0;
                off += vsz;
                slow = 1;
            }
        }
    }
    slot->nvalid = attnum;
    slot->off = off;
    slot->slow = slow;
}

static void store_col(int type, void *col, int64_t i, int64_t v)
{
    switch (type) {
        case GX_INT4:
        case GX_DATE:   ((int32_t *) col)[i] = (int32_t) v; break;
        case GX_INT8:
        case GX_FLOAT8: ((int64_t *) col)[i] = v; break;
        case GX_CHAR:
        case ORC_BPCHAR1: ((int8_t *) col)[i] = (int8_t) v; break;
    }
}

int64_t orc_rel_scan_columns(const orc_rel *r, int ncols, const int32_t *attnums,
                             void *const *cols_out, uint8_t *const *nulls_out)
{
    int64_t values[64];
    uint8_t isnull[64];
    uint16_t vis[ORC_BLCKSZ / 4];
    orc_slot slot;
    int64_t n = 0;
    int maxatt = 0;
    for (int c = 0; c < ncols; c++) if (attnums[c] + 1 > maxatt) maxatt = attnums[c] + 1;
    memset(&slot, 0, sizeof(slot));
    slot.natts = r->natts; slot.attrs = r->attrs; slot.values = values; slot.isnull = isnull;
    for (int64_t p = 0; p < r->npages; p++) {
        const uint8_t *pg = r->pages[p];
        int nv = orc_heapgetpage(pg, vis);
        for (int k = 0; k < nv; k++) {
            uint32_t lp = rd32(pg + ORC_PAGE_HDR + 4 * (uint32_t) (vis[k] - 1));
            slot.tuple = pg + (lp & 0x7FFF);
            slot.nvalid = 0; slot.off = 0; slot.slow = 0;
            orc_slot_deform(&slot, maxatt);
            for (int c = 0; c < ncols; c++) {
                int a = attnums[c];
                if (nulls_out && nulls_out[c]) nulls_out[c][n] = isnull[a];
                store_col(r->attrs[a].type, cols_out[c], n, isnull[a] ? 0 : values[a]);
            }
            n++;
        }
    }
    return n;
}
