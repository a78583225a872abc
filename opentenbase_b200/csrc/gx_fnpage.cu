// gx_fnpage.cu — forward-node pages: the reference's redistribute wire format, written and read on the device.
//
// Between datanodes the reference moves tuples in 8 KB "FnPages" (forward/fnbufpage.h:54-65): a 32-byte header
// {lower, fid, nodeid, queryid, flag, workerid, virtualid} followed by MAXALIGNed MINIMAL tuples back to back; the
// sender forms each tuple in place (FragmentSendAttrs, executor/execFragment.c:2067-2136 ->
// heap_form_minimal_tuple_ptr, access/common/heaptuple.c:1852-1895), starts a new page when the aligned tuple no longer
// fits (FragmentGetPage, execFragment.c:1857-1876) and ends a stream with a MAX_UINT32 length word + FNPAGE_END
// (FragmentSendNullTuple, :1963-1975); the receiver walks the (length, data) chain (fnbufpage.h:107-127,
// executor/tqueueThread.c:913-925) and deforms each tuple (slot_deform_tuple, heaptuple.c:1518-1614).
// A GPU datanode that talks to stock CPU datanodes through their forwarder needs exactly these two conversions:
//   gx_fnpage_pack    columns in HBM  -> pages (one warp per page; byte-identical to the reference's pages when no
//                                        column carries NULLs, see below)
//   gx_fnpage_unpack  pages           -> columns in HBM (one warp per page: the page is staged in shared memory, one
//                                        lane walks the chain, all lanes deform)
// Both are checked against oracle/orc_fnpage.c, which is pinned byte for byte to the reference's heaptuple.o and
// fnbufpage.o (tests/golden/fnpage_vectors.json).
#include "gx_internal.cuh"

#define FNP_BLCKSZ     8192
#define FNP_HDR        32
#define FNP_END        8u
#define FNP_HUGE       1u
#define MT_HDR         15          /* SizeofMinimalTupleHeader (htup_details.h:784 with _PG_ORCL_ and _SHARDING_) */
#define MT_OFFSET      32          /* MINIMAL_TUPLE_OFFSET */
#define MT_INVALID_SHARD 4096      /* InvalidShardID, postgres_ext.h:80-81 */
#define HEAP_HASNULL_BIT     0x0001
#define HEAP_HASVARWIDTH_BIT 0x0002

struct gx_fnp_cols {
    int natts;
    const void *data[GX_MAX_COLS]; const uint8_t *nulls[GX_MAX_COLS];
    short att_len[GX_MAX_COLS]; signed char att_align[GX_MAX_COLS];
};
struct gx_fnp_pack_args {
    gx_fnp_cols c;
    long long nrows, rows_per_page, npages_data;
    int fixed_len;                 // aligned tuple size when no column has a NULL array, else 0
    long long end_page;            // page that carries the end-of-stream word (-1: none)
    gx_fnpage_id id;
    uint8_t *pages;
};

__device__ __forceinline__ long long fnp_load(const void *col, int len, long long r)
{
    switch (len) {
        case 8: return ((const long long *) col)[r];
        case 4: return (long long) ((const int *) col)[r];
        default: return (long long) ((const signed char *) col)[r];      // 1-byte "char" and bpchar(1) (len -1)
    }
}

// one warp per page
__global__ void __launch_bounds__(256) gx_k_fnpage_pack(const __grid_constant__ gx_fnp_pack_args a)
{
    const long long p = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long npages = a.end_page >= a.npages_data ? a.end_page + 1 : a.npages_data;
    if (p >= npages) return;
    uint8_t *pg = a.pages + p * FNP_BLCKSZ;
    const long long r0 = p * a.rows_per_page;
    long long cnt = p < a.npages_data ? a.nrows - r0 : 0;
    if (cnt > a.rows_per_page) cnt = a.rows_per_page;
    unsigned int lower = FNP_HDR;
    for (long long j0 = 0; j0 < cnt; j0 += 32) {
        const long long j = j0 + lane, r = r0 + j;
        const bool act = j < cnt;
        // ---- size of this lane's tuple: heap_minimal_tuple_header_size + heap_compute_data_size
        unsigned int hoff = 0, len = 0, alen = 0; bool hasnull = false, hasvar = false;
        unsigned long long nullmask = 0;
        if (act) {
            for (int i = 0; i < a.c.natts; i++) if (a.c.nulls[i] && a.c.nulls[i][r]) nullmask |= 1ULL << i;
            hasnull = nullmask != 0;
            hoff = (MT_HDR + (hasnull ? (a.c.natts + 7) / 8 : 0) + 7) & ~7u;
            unsigned int off = 0;
            for (int i = 0; i < a.c.natts; i++) {
                if ((nullmask >> i) & 1ULL) continue;
                if (a.c.att_len[i] == -1) { off += 2; hasvar = true; }                   // short varlena: 1-byte header + the character
                else { off = (off + a.c.att_align[i] - 1) & ~(unsigned) (a.c.att_align[i] - 1); off += a.c.att_len[i]; }
            }
            len = hoff + off; alen = (len + 7) & ~7u;
        }
        // ---- where it goes: tuples of a page lie back to back in row order
        unsigned int inc = alen;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        const unsigned int at = lower + inc - alen;
        lower += __shfl_sync(0xffffffffu, inc, 31);
        if (!act) continue;
        uint8_t *t = pg + at;                                                            // 8-aligned; the buffer was zeroed
        // {t_len u32, pad u16, t_infomask2 u16} {t_infomask u16, t_infomask3 u16, t_shardid u16, t_hoff u8, t_bits[0] u8}
        const unsigned int infomask = (hasnull ? HEAP_HASNULL_BIT : 0) | (hasvar ? HEAP_HASVARWIDTH_BIT : 0);
        unsigned long long w0 = (unsigned long long) len | ((unsigned long long) (a.c.natts & 0x07FF) << 48);
        unsigned long long bits = 0;                                                     // att present <=> bit set
        if (hasnull) bits = ~nullmask & (a.c.natts >= 64 ? ~0ULL : ((1ULL << a.c.natts) - 1));
        unsigned long long w1 = (unsigned long long) infomask | ((unsigned long long) MT_INVALID_SHARD << 32) |
                                ((unsigned long long) (hoff + MT_OFFSET) << 48) | ((bits & 0xffULL) << 56);
        *(unsigned long long *) t = w0;
        *(unsigned long long *) (t + 8) = w1;
        if (hasnull) for (int b = 1; b < (a.c.natts + 7) / 8; b++) t[MT_HDR + b] = (uint8_t) (bits >> (8 * b));
        uint8_t *d = t + hoff;
        unsigned int off = 0;
        for (int i = 0; i < a.c.natts; i++) {                                            // heap_fill_tuple
            if ((nullmask >> i) & 1ULL) continue;
            const int al = a.c.att_len[i];
            const long long v = fnp_load(a.c.data[i], al, r);
            if (al == -1) { d[off] = (uint8_t) ((2 << 1) | 1); d[off + 1] = (uint8_t) v; off += 2; continue; }
            off = (off + a.c.att_align[i] - 1) & ~(unsigned) (a.c.att_align[i] - 1);
            if (al == 8) *(long long *) (d + off) = v;
            else if (al == 4) *(int *) (d + off) = (int) v;
            else d[off] = (uint8_t) v;
            off += al;
        }
    }
    if (lane == 0) {
        unsigned int flag = 0;
        if (p == a.end_page) { *(unsigned int *) (pg + lower) = 0xffffffffu; lower += 4; flag = FNP_END; }
        // FnPageInit + FragmentGetPage: {lower u32, fid u16, nodeid u16} {qid} {qid} {flag u32, workerid u16, virtualid u8, pad u8}
        *(unsigned long long *) pg = (unsigned long long) lower | ((unsigned long long) a.id.fid << 32) | ((unsigned long long) a.id.nodeid << 48);
        *(long long *) (pg + 8) = a.id.qid_timestamp_nodeid;
        *(long long *) (pg + 16) = a.id.qid_sequence;
        *(unsigned long long *) (pg + 24) = (unsigned long long) flag | ((unsigned long long) a.id.workerid << 32) | ((unsigned long long) a.id.virtualid << 48);
    }
}

static int fnp_check_desc(gx_ctx *ctx, const gx_heap_desc *desc, int ncols, const int32_t *types, const char *who)
{
    GX_CHECK_ARG(ctx, desc->natts >= 1 && desc->natts <= GX_MAX_COLS && desc->ncols == ncols, "%s: descriptor has %d attributes / %d columns, table %d (at most %d)", who, desc->natts, desc->ncols, ncols, GX_MAX_COLS);
    for (int c = 0; c < ncols; c++) {
        const int a = desc->attnums[c];
        GX_CHECK_ARG(ctx, a >= 0 && a < desc->natts, "%s: column %d maps to attribute %d", who, c, a);
        const int want = gx_type_size(types[c]), al = desc->att_len[a];
        GX_CHECK_ARG(ctx, al == want || (al == -1 && types[c] == GX_CHAR), "%s: column %d (type %d) against attlen %d", who, c, types[c], al);
    }
    for (int a = 0; a < desc->natts; a++) {
        const int al = desc->att_len[a], ag = desc->att_align[a];
        GX_CHECK_ARG(ctx, (al == 1 || al == 2 || al == 4 || al == 8 || al == -1) && (ag == 1 || ag == 2 || ag == 4 || ag == 8), "%s: attribute %d: attlen %d attalign %d", who, a, al, ag);
    }
    return GX_OK;
}

extern "C" int gx_fnpage_pack(gx_ctx *ctx, const gx_table *t, const gx_heap_desc *desc, const gx_fnpage_id *id, int end_marker,
                              void *host_pages, int64_t cap_pages, int64_t *npages_out)
{
    if (!ctx || !t || !desc || !id || !npages_out) return GX_ERR_ARG;
    int rc = fnp_check_desc(ctx, desc, t->ncols, t->types, "fnpage_pack"); if (rc) return rc;
    GX_CHECK_ARG(ctx, desc->natts == t->ncols, "fnpage_pack: the tuple has %d attributes, the table %d columns", desc->natts, t->ncols);
    gx_fnp_pack_args a; memset(&a, 0, sizeof(a));
    a.c.natts = desc->natts; a.nrows = t->nrows; a.id = *id; a.id._pad = 0;
    bool anynull = false;
    unsigned int off = 0;
    for (int c = 0; c < t->ncols; c++) {
        GX_CHECK_ARG(ctx, desc->attnums[c] == c, "fnpage_pack: column %d must feed attribute %d", c, c);
        GX_CHECK_ARG(ctx, desc->att_len[c] != 2, "fnpage_pack: 2-byte attributes have no column type here");
        a.c.data[c] = t->cols[c]; a.c.nulls[c] = t->nulls[c]; a.c.att_len[c] = desc->att_len[c]; a.c.att_align[c] = desc->att_align[c];
        anynull |= t->nulls[c] != nullptr;
        if (desc->att_len[c] == -1) off += 2;
        else { off = (off + desc->att_align[c] - 1) & ~(unsigned) (desc->att_align[c] - 1); off += desc->att_len[c]; }
    }
    // the widest tuple: every attribute present (dropping one never lengthens the walk) and, if NULLs can occur, the bitmap
    const unsigned int hoff_max = (MT_HDR + (anynull ? (desc->natts + 7) / 8 : 0) + 7) & ~7u;
    const unsigned int lmax = (hoff_max + off + 7) & ~7u;
    GX_CHECK_ARG(ctx, lmax <= FNP_BLCKSZ - FNP_HDR, "fnpage_pack: a %u-byte tuple needs FragmentSendHuge", lmax);
    // Without NULL arrays every tuple has this size and the reference's greedy fill is "floor(8160 / size) per page": the
    // pages come out byte-identical.  With NULLs sizes vary per row and greedy page breaks form a sequential chain; rows
    // per page are then fixed at what fits in the worst case - valid pages, a little emptier than the reference's.
    a.fixed_len = anynull ? 0 : (int) lmax;
    a.rows_per_page = (FNP_BLCKSZ - FNP_HDR) / lmax;
    a.npages_data = (t->nrows + a.rows_per_page - 1) / a.rows_per_page;
    a.end_page = -1;
    if (end_marker) {
        bool room = false;                                         // FragmentSendNullTuple: 4 bytes on the current page, else a new one
        if (!anynull && a.npages_data > 0) {
            const long long last = t->nrows - (a.npages_data - 1) * a.rows_per_page;
            room = FNP_BLCKSZ - (FNP_HDR + last * lmax) >= 4;
        }
        a.end_page = room ? a.npages_data - 1 : a.npages_data;
    }
    const int64_t npages = a.end_page >= a.npages_data ? a.end_page + 1 : a.npages_data;
    *npages_out = npages;
    if (!host_pages || npages == 0) return GX_OK;
    GX_CHECK_ARG(ctx, cap_pages >= npages, "fnpage_pack: %lld pages do not fit the caller's %lld", (long long) npages, (long long) cap_pages);
    uint8_t *d_pages;
    cudaError_t e = gx_tmp_alloc(ctx, (void **) &d_pages, (size_t) npages * FNP_BLCKSZ);
    if (e != cudaSuccess) { GX_SET_ERR(ctx, "fnpage_pack: %s", cudaGetErrorString(e)); return GX_ERR_NOMEM; }
    a.pages = d_pages;
    // every byte that leaves the node is defined: nothing of whatever the pool block held before goes onto the wire
    e = cudaMemsetAsync(d_pages, 0, (size_t) npages * FNP_BLCKSZ, ctx->stream);
    if (e == cudaSuccess) {
        gx_launch_scope ls(ctx, "fnpage_pack");
        gx_k_fnpage_pack<<<(unsigned) ((npages * 32 + 255) / 256), 256, 0, ctx->stream>>>(a);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_pages, d_pages, (size_t) npages * FNP_BLCKSZ, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    gx_tmp_free(ctx, d_pages);
    if (e != cudaSuccess) { GX_SET_ERR(ctx, "fnpage_pack: %s", cudaGetErrorString(e)); return GX_ERR_CUDA; }
    return GX_OK;
}

// ----------------------------------------------------------------------------- receiver
struct gx_fnp_unpack_args {
    const uint8_t *pages; long long npages;
    int natts, ncols;
    short att_len[64]; signed char att_align[64]; signed char col_of_att[64];
    void *out[GX_MAX_COLS]; uint8_t *out_nulls[GX_MAX_COLS]; int out_type[GX_MAX_COLS];
    long long *counts;             // per page: tuples (count pass) / first row (deform pass, after the scan)
    int *error;                    // 1 FNPAGE_HUGE, 2 a length word that leaves the page, 4 NULL in a column without a NULL array
};
#define FNU_WARPS 4
#define FNU_MAXT  ((FNP_BLCKSZ - FNP_HDR) / 16)     /* a tuple takes at least 16 bytes */

// DEFORM false: count the tuples of every page; true: walk again and store the attributes
template <bool DEFORM>
__global__ void __launch_bounds__(FNU_WARPS * 32) gx_k_fnpage_unpack(const __grid_constant__ gx_fnp_unpack_args a)
{
    __shared__ __align__(16) uint8_t s_page[FNU_WARPS][FNP_BLCKSZ];
    __shared__ unsigned short s_off[FNU_WARPS][FNU_MAXT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long p = (long long) blockIdx.x * FNU_WARPS + warp; p < a.npages; p += (long long) gridDim.x * FNU_WARPS) {
        const uint8_t *g = a.pages + p * FNP_BLCKSZ;
        uint8_t *pg = s_page[warp];
        const unsigned int lower_g = *(const unsigned int *) g;
        const unsigned int used = lower_g <= FNP_BLCKSZ ? lower_g : FNP_BLCKSZ;
        for (unsigned int i = lane * 16; i < used; i += 32 * 16) *(uint4 *) (pg + i) = *(const uint4 *) (g + i);
        __syncwarp();
        int n = 0;
        if (lane == 0) {
            if (*(const unsigned int *) (pg + 24) & FNP_HUGE) atomicOr(a.error, 1);
            else if (lower_g > FNP_BLCKSZ) atomicOr(a.error, 2);
            else {
                unsigned int off = FNP_HDR;                                     // InitFnPageIterator
                while (off < lower_g) {                                         // FnPageIterateDone
                    const unsigned int len = *(const unsigned int *) (pg + off);
                    if (len == 0xffffffffu) break;                              // end of stream
                    if (len < 16 || off + len > lower_g) { atomicOr(a.error, 2); break; }
                    s_off[warp][n++] = (unsigned short) off;
                    off += (len + 7) & ~7u;                                     // FnPageIterateNext
                }
            }
        }
        n = __shfl_sync(0xffffffffu, n, 0);
        if (!DEFORM) { if (lane == 0) a.counts[p] = n; __syncwarp(); continue; }
        const long long row0 = a.counts[p];
        for (int j = lane; j < n; j += 32) {
            const uint8_t *t = pg + s_off[warp][j];
            const long long row = row0 + j;
            const bool hasnulls = (*(const unsigned short *) (t + 8) & HEAP_HASNULL_BIT) != 0;
            const int tnatts = (int) (*(const unsigned short *) (t + 6) & 0x07FF);
            const uint8_t *bp = t + MT_HDR;
            const uint8_t *tp = t + t[14] - MT_OFFSET;
            unsigned int off = 0;
            for (int att = 0; att < a.natts; att++) {                           // slot_deform_tuple
                const int c = a.col_of_att[att];
                if (att >= tnatts || (hasnulls && !(bp[att >> 3] & (1 << (att & 7))))) {
                    if (c >= 0) {
                        if (a.out_nulls[c]) a.out_nulls[c][row] = 1; else atomicOr(a.error, 4);
                        switch (a.out_type[c]) { case GX_INT4: case GX_DATE: ((int *) a.out[c])[row] = 0; break;
                                                 case GX_CHAR: ((signed char *) a.out[c])[row] = 0; break;
                                                 default: ((long long *) a.out[c])[row] = 0; }
                    }
                    continue;
                }
                const int len = a.att_len[att], al = a.att_align[att];
                long long v = 0;
                if (len > 0) {
                    off = (off + al - 1) & ~(unsigned) (al - 1);                // att_align_nominal
                    switch (len) {
                        case 1: v = (long long) (signed char) tp[off]; break;
                        case 2: v = (long long) *(const short *) (tp + off); break;
                        case 4: v = (long long) *(const int *) (tp + off); break;
                        default: v = *(const long long *) (tp + off); break;
                    }
                    off += len;
                } else {
                    if (tp[off] == 0) off = (off + al - 1) & ~(unsigned) (al - 1);   // att_align_pointer
                    const uint8_t h = tp[off];
                    unsigned int vsz, hdr;
                    if (h & 1) { vsz = (h >> 1) & 0x7F; hdr = 1; }
                    else { vsz = ((((unsigned) tp[off]) | ((unsigned) tp[off + 1] << 8) | ((unsigned) tp[off + 2] << 16) | ((unsigned) tp[off + 3] << 24)) >> 2) & 0x3FFFFFFF; hdr = 4; }
                    v = vsz > hdr ? (long long) (signed char) tp[off + hdr] : 0;
                    off += vsz;
                }
                if (c >= 0) {
                    if (a.out_nulls[c]) a.out_nulls[c][row] = 0;
                    switch (a.out_type[c]) { case GX_INT4: case GX_DATE: ((int *) a.out[c])[row] = (int) v; break;
                                             case GX_CHAR: ((signed char *) a.out[c])[row] = (signed char) v; break;
                                             default: ((long long *) a.out[c])[row] = v; }
                }
            }
        }
        __syncwarp();
    }
}

__global__ void gx_k_scan_inplace(long long *v, long long n, long long *total);   // gx_table.cu

extern "C" int gx_fnpage_unpack(gx_ctx *ctx, const void *host_pages, int64_t npages, const gx_heap_desc *desc,
                                const int32_t *col_types, gx_table **out)
{
    if (!ctx || !desc || !col_types || !out || npages < 0 || (npages > 0 && !host_pages)) return GX_ERR_ARG;
    GX_CHECK_ARG(ctx, desc->natts >= 1 && desc->natts <= 64 && desc->ncols >= 1 && desc->ncols <= GX_MAX_COLS, "fnpage_unpack: %d attributes, %d columns", desc->natts, desc->ncols);
    gx_fnp_unpack_args a; memset(&a, 0, sizeof(a));
    a.npages = npages; a.natts = desc->natts; a.ncols = desc->ncols;
    bool hn[GX_MAX_COLS];
    for (int i = 0; i < desc->natts; i++) {
        const int al = desc->att_len[i], ag = desc->att_align[i];
        GX_CHECK_ARG(ctx, (al == 1 || al == 2 || al == 4 || al == 8 || al == -1) && (ag == 1 || ag == 2 || ag == 4 || ag == 8), "fnpage_unpack: attribute %d: attlen %d attalign %d", i, al, ag);
        a.att_len[i] = (short) al; a.att_align[i] = (signed char) ag; a.col_of_att[i] = -1;
    }
    for (int c = 0; c < desc->ncols; c++) {
        const int at = desc->attnums[c];
        GX_CHECK_ARG(ctx, at >= 0 && at < desc->natts && a.col_of_att[at] < 0, "fnpage_unpack: column %d maps to attribute %d", c, at);
        const int al = desc->att_len[at], want = gx_type_size(col_types[c]);
        GX_CHECK_ARG(ctx, want > 0 && (al == want || (al == -1 && col_types[c] == GX_CHAR)), "fnpage_unpack: column %d (type %d) against attlen %d", c, col_types[c], al);
        a.col_of_att[at] = (signed char) c; a.out_type[c] = col_types[c];
        hn[c] = !desc->att_notnull[at];
    }
    uint8_t *d_pages = nullptr; long long *d_counts = nullptr; int *d_err = nullptr;
    cudaError_t e = gx_tmp_alloc(ctx, (void **) &d_pages, (size_t) (npages > 0 ? npages : 1) * FNP_BLCKSZ);
    if (e == cudaSuccess) e = gx_tmp_alloc(ctx, (void **) &d_counts, (size_t) (npages + 2) * sizeof(long long));
    if (e != cudaSuccess) { gx_tmp_free(ctx, d_pages); GX_SET_ERR(ctx, "fnpage_unpack: %s", cudaGetErrorString(e)); return GX_ERR_NOMEM; }
    d_err = (int *) (d_counts + npages + 1);
    a.pages = d_pages; a.counts = d_counts; a.error = d_err;
    long long total = 0; int h_err = 0;
    const unsigned grid = (unsigned) ((npages + FNU_WARPS - 1) / FNU_WARPS < (long long) ctx->sm_count * 6 ? (npages + FNU_WARPS - 1) / FNU_WARPS : (long long) ctx->sm_count * 6);
    e = cudaMemsetAsync(d_err, 0, sizeof(long long), ctx->stream);
    if (e == cudaSuccess && npages > 0) e = cudaMemcpyAsync(d_pages, host_pages, (size_t) npages * FNP_BLCKSZ, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && npages > 0) {
        gx_launch_scope ls(ctx, "fnpage_unpack", 2);
        gx_k_fnpage_unpack<false><<<grid, FNU_WARPS * 32, 0, ctx->stream>>>(a);
        gx_k_scan_inplace<<<1, 1024, 0, ctx->stream>>>(d_counts, npages, ctx->d_scratch);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && npages > 0) total = ctx->h_scratch[0];
    int rc = GX_OK;
    gx_table *t = nullptr;
    if (e == cudaSuccess && h_err) { GX_SET_ERR(ctx, "fnpage_unpack: %s", (h_err & 1) ? "a page carries FNPAGE_HUGE (tuples above 8160 bytes are not handled)" : "a length word points outside its page"); rc = GX_ERR_ARG; }
    if (e == cudaSuccess && rc == GX_OK) rc = gx_table_alloc_like(ctx, desc->ncols, col_types, hn, total > 0 ? total : 1, &t);
    if (e == cudaSuccess && rc == GX_OK && total > 0) {
        for (int c = 0; c < desc->ncols; c++) { a.out[c] = t->cols[c]; a.out_nulls[c] = t->nulls[c]; }
        { gx_launch_scope ls(ctx, "fnpage_unpack"); gx_k_fnpage_unpack<true><<<grid, FNU_WARPS * 32, 0, ctx->stream>>>(a); }
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess && (h_err & 4)) { GX_SET_ERR(ctx, "fnpage_unpack: NULL in a column declared NOT NULL"); rc = GX_ERR_STATE; }
    }
    gx_tmp_free(ctx, d_pages); gx_tmp_free(ctx, d_counts);
    if (e != cudaSuccess) { if (t) gx_table_free(t); GX_SET_ERR(ctx, "fnpage_unpack: %s", cudaGetErrorString(e)); return GX_ERR_CUDA; }
    if (rc != GX_OK) { if (t) gx_table_free(t); return rc; }
    t->nrows = total;
    *out = t;
    return GX_OK;
}
