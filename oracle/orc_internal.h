/* orc_internal.h — shared between the oracle's translation units.
 * TEST INFRASTRUCTURE (see otb_oracle.h). */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include "otb_oracle.h"

#define ORC_BLCKSZ            8192
#define ORC_MAXALIGN(x)       (((uintptr_t) (x) + 7) & ~(uintptr_t) 7)
#define ORC_PAGE_HDR          44     /* offsetof(PageHeaderData, pd_linp), bufpage.h:153-175: LocationIndex is
                                      * uint32 under __OPENTENBASE_C__ (bufpage.h:85-89); pinned by oracle/_ref */
#define ORC_HEAP_HDR          47     /* offsetof(HeapTupleHeaderData, t_bits), htup_details.h:165-201 */
#define ORC_MINIMAL_TUPLE_OFFSET 32  /* ((38 - 4) / 8) * 8, htup_details.h:743-744 */
#define ORC_BPCHAR1           6      /* heap-side only: bpchar(1), short varlena */

/* offsets inside HeapTupleHeaderData (htup_details.h:126-201 with
 * __SUPPORT_DISTRIBUTED_TRANSACTION__, _PG_ORCL_, _SHARDING_ all on) */
#define HTH_XMIN       0
#define HTH_XMAX       4
#define HTH_XMAX_TS    8
#define HTH_XMIN_TS    16
#define HTH_CID        24
#define HTH_CTID       32
#define HTH_INFOMASK2  38
#define HTH_INFOMASK   40
#define HTH_INFOMASK3  42
#define HTH_SHARDID    44
#define HTH_HOFF       46
#define HTH_BITS       47

#define HEAP_HASNULL         0x0001
#define HEAP_HASVARWIDTH     0x0002
#define HEAP_XMIN_COMMITTED  0x0100
#define HEAP_XMAX_COMMITTED  0x0400
#define HEAP_XMAX_INVALID    0x0800
#define HEAP_NATTS_MASK      0x07FF

#define LP_UNUSED 0
#define LP_NORMAL 1

typedef struct orc_attr {
    int32_t type;      /* GX_* or ORC_BPCHAR1 */
    int16_t attlen;    /* 1,2,4,8 or -1 */
    int8_t  attalign;  /* 1,2,4,8 */
    int32_t attcacheoff;
} orc_attr;

struct orc_rel {
    int       natts;
    orc_attr *attrs;
    int64_t   npages, pages_cap;
    uint8_t **pages;          /* each ORC_BLCKSZ, 8-aligned */
    int64_t   ntuples;
};

/* a TupleTableSlot stand-in (tuptable.h:181-216) */
typedef struct orc_slot {
    int       natts;
    int       nvalid;          /* tts_nvalid */
    uint32_t  off;             /* tts_off */
    int       slow;            /* TTS_FLAG_SLOW */
    int       empty;           /* TTS_FLAG_EMPTY */
    const uint8_t *tuple;      /* HeapTupleHeader, or NULL if virtual */
    orc_attr *attrs;           /* tuple descriptor */
    int64_t  *values;          /* tts_values (Datum) */
    uint8_t  *isnull;          /* tts_isnull */
} orc_slot;

void orc_type_layout(int type, int16_t *attlen, int8_t *attalign);
int  orc_heapgetpage(const uint8_t *pg, uint16_t *vistuples);   /* heapam.c:388 */
void orc_slot_deform(orc_slot *slot, int natts);           /* slot_deform_tuple */
static inline int64_t orc_slot_getattr(orc_slot *slot, int attnum /* 0-based */, uint8_t *isnull)
{
    /* slot_getattr, heaptuple.c:1630-1700 */
    if (attnum >= slot->nvalid && slot->tuple)
        orc_slot_deform(slot, attnum + 1);
    *isnull = slot->isnull[attnum];
    return slot->values[attnum];
}

/* heap_compute_data_size + heap_fill_tuple body only (no header):
 * forms `natts` datums at dst (8-aligned), returns data length; *hasnull set */
uint32_t orc_form_data(const orc_attr *attrs, int natts, const int64_t *values,
                       const uint8_t *isnull, uint8_t *bits /* may be NULL */,
                       uint8_t *dst);
uint32_t orc_compute_data_size(const orc_attr *attrs, int natts, const int64_t *values,
                               const uint8_t *isnull);

#endif
