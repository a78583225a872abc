// gx_internal.cuh — internal structures and device helpers of libgpuexec.so.
// sm_100a only.  No CPU fallback anywhere: every entry point needs a live ctx.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <map>
#include "../../include/gpuexec.h"

#define GX_EMPTY_KEY   ((long long) 0x8000000000000000LL)   /* INT64_MIN marks an empty join slot */
#define GX_MAX_NODES   64
#define GX_SUB_LOG2    11                 /* join table: linear probing wraps inside 2048-slot (32 KB) sub-tables */
#define GX_SUB         (1 << GX_SUB_LOG2)
#define GX_NCOPY       4                  /* copy streams of the host-table loader */

// ---------------------------------------------------------------- host side
struct gx_prof_entry { double ms; int64_t launches; };
#define GX_PROF_POOL 2048
struct gx_prof_rec { cudaEvent_t a, b; const char *name; };

struct gx_nccl_api;   // gx_comm.cu

struct gx_ctx {
    int device;
    int sm_count, cc_major, cc_minor;
    size_t hbm_bytes;
    size_t smem_optin;           // max dynamic smem per block
    cudaStream_t stream;         // all kernels
    cudaStream_t copy_streams[GX_NCOPY];   // H2D staging: chunks go round-robin over these (gx_exec.cu)
    cudaEvent_t copy_ev[2][GX_NCOPY];      // [0] inner table landed, [1] outer table landed (per copy stream)
    cudaEvent_t ev_alloc;                  // the copy streams wait for the stream-ordered allocations
    cudaEvent_t ev_t0, ev_t1;    // gx_timer_*
    char err[512];
    int64_t launches;
    int profile;
    std::map<std::string, gx_prof_entry> *prof;
    gx_prof_rec *prof_pool; int prof_used;      // events recorded, not yet resolved
    void *l2flush_buf; size_t l2flush_bytes;
    // pinned staging ring of the heap-page loader (gx_stage_acquire): the host fills one slot while the
    // DMA of the other is in flight
    void *stage[2]; size_t stage_bytes[2]; cudaEvent_t stage_ev[2]; int stage_busy[2]; int stage_next;
    // small device scratch (counters / flags)
    long long *d_scratch;        // 64 x int64
    long long *h_scratch;        // pinned mirror
    // routing
    int32_t *d_shardmap; int nnodes;
    // communicator
    gx_nccl_api *nccl; void *comm; int rank, nranks;
    // peer windows (gx_comm.cu): every rank's exchange block mapped into this process over NVLink;
    // peer_base[p] = rank p's block (control words, then the data window), peer_ready 1 = mapped, -1 = not available
    int peer_ready; size_t peer_win_bytes; char *peer_base[GX_MAX_NODES]; unsigned long long peer_epoch;
};

struct gx_table {
    gx_ctx *ctx;
    int ncols;
    int32_t types[GX_MAX_COLS];
    void *cols[GX_MAX_COLS];
    uint8_t *nulls[GX_MAX_COLS];     // NULL when the column has no NULLs
    int64_t nrows, capacity;
};

struct gx_slot { long long key; unsigned long long payload; };   // 16 B
// Compact slot (8 B).  With an order-preserving slot function the keys that can land in one
// 2048-slot sub-table lie within a few thousand key units of each other, so 31 bits of
// (key - kmin) identify a key inside its sub-table whatever the total key span is:
// d = ((key - kmin) << 1) | 1 truncated to 32 bits; d is never 0, which marks an empty slot.
// Half the bytes to write, read and keep in L2.  Needs a payload of at most 4 bytes and the
// exact kmin/kmax of the build side (probe keys outside [kmin, kmax] are rejected before the
// 31-bit compare).
#define GX_CSLOT_D(key, kmin) ((unsigned int) (((((unsigned long long) (key)) - ((unsigned long long) (kmin))) << 1) | 1ULL))
struct gx_cslot { unsigned int d; unsigned int payload; };

struct gx_hash {
    gx_ctx *ctx;
    gx_slot *slots;
    int64_t nslots;                 // power of two
    int64_t nentries;
    int key_type;
    int n_payload;                  // 0: payload = build row number
    int32_t payload_types[GX_MAX_PAYLOAD];
    int unique;
    int mode;                       // slot function: 0 mixing hash, 1 order-preserving interpolation (gx_slot_index)
    long long kmin; unsigned long long scale; unsigned int win; unsigned int shift;
    double avg_chain;               // measured while the table was filled
    int sorted_build;               // built by the partition-free key-ordered path
    // compact form (gx_cslot): produced by the key-ordered build when key range and payload allow;
    // the 16-byte form is materialised from it on demand for consumers that only read that one
    gx_cslot *cslots; unsigned long long cspan;   // cspan = kmax - kmin + 1
    double keys_per_slot;           // key range / slot range of the interpolation (compact -> wide reconstruction)
    unsigned int amask;             // home slots are aligned to amask + 1 (2 for 16 B slots, 4 for compact ones)
    // rows whose key equals GX_EMPTY_KEY cannot live in the table: side list
    unsigned long long *special_payload; int special_cap; int special_count;
};

int gx_hash_wide(gx_ctx *ctx, gx_hash *h);
cudaError_t gx_stage_mark(gx_ctx *ctx, const void *host);   // gx_exec.cu: record "copy of this staging slot enqueued"   // materialise the 16-byte slot array of a compact table

// A group-state record: [k0][k1][w0..w(nwords-1)], 8-byte words.
// w0 is always the group's row count.
#define GX_MAX_WORDS 18

struct gx_result {
    gx_ctx *ctx;
    gx_agg_plan plan;
    int32_t group_types[GX_MAX_GROUP_COLS];
    int nkw;                         // key words (1 or 2)
    int nwords;                      // state words
    int rec_words;                   // 2 + nwords (k0,k1 always present in records)
    // per-agg state layout
    int agg_word[GX_MAX_AGGS];       // first word of the aggregate's state
    int agg_cnt_word[GX_MAX_AGGS];   // word holding its non-NULL input count (may be 0 = row count)
    long long *d_recs;               // ngroups * rec_words (dense, partial states)
    unsigned int *d_nullmask;        // ngroups (group-key null bits)
    int64_t ngroups, cap;
    int finalized_across;            // combined over the communicator already
    int need_w0;                     // w0 (rows per group) is maintained
};

#define GX_SET_ERR(ctx, ...) do { if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
                                  gx_set_global_err(__VA_ARGS__); } while (0)
void gx_set_global_err(const char *fmt, ...);

#define GX_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    GX_SET_ERR(ctx, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
    return (e__ == cudaErrorMemoryAllocation) ? GX_ERR_NOMEM : GX_ERR_CUDA; } } while (0)

#define GX_CHECK_ARG(ctx, cond, ...) do { if (!(cond)) { GX_SET_ERR(ctx, __VA_ARGS__); return GX_ERR_ARG; } } while (0)

// launch accounting + optional per-kernel CUDA-event timing
struct gx_launch_scope {
    gx_ctx *ctx; const char *name;
    int slot;
    gx_launch_scope(gx_ctx *c, const char *n, int nlaunch = 1) : ctx(c), name(n), slot(-1) {
        ctx->launches += nlaunch;
        // events are only recorded here (asynchronously); gx_profile_get() resolves them
        if (ctx->profile && ctx->prof_used < GX_PROF_POOL) {
            slot = ctx->prof_used++;
            ctx->prof_pool[slot].name = name;
            cudaEventRecord(ctx->prof_pool[slot].a, ctx->stream);
        }
    }
    ~gx_launch_scope() { if (slot >= 0) cudaEventRecord(ctx->prof_pool[slot].b, ctx->stream); }
};

static inline int gx_type_size(int t)
{
    switch (t) { case GX_INT4: case GX_DATE: return 4; case GX_INT8: case GX_FLOAT8: return 8; case GX_CHAR: return 1; }
    return 0;
}
static inline int64_t gx_pow2_ceil(int64_t x) { int64_t p = 1; while (p < x) p <<= 1; return p; }

// Stream-ordered, pool-cached device allocations for per-query temporaries (join table,
// bucket buffers, group tables): cudaMalloc/cudaFree of multi-GB blocks cost milliseconds per query.
static inline cudaError_t gx_tmp_alloc(gx_ctx *ctx, void **p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 8, ctx->stream); }
static inline void gx_tmp_free(gx_ctx *ctx, void *p) { if (p) cudaFreeAsync(p, ctx->stream); }

int gx_table_alloc_like(gx_ctx *ctx, int ncols, const int32_t *types, const bool *has_nulls,
                        int64_t capacity, gx_table **out);
int gx_result_alloc(gx_ctx *ctx, const gx_agg_plan *plan, const int32_t *group_types, int64_t cap, gx_result **out);

// device-side plan descriptors -------------------------------------------
struct gx_dcol { const void *data; const uint8_t *nulls; int type; int _pad; };

struct gx_dpred { gx_dcol col; int op; int _pad; long long ival; double fval; };

// chain-form expression: acc = term0; acc = acc (op) term_i
enum { GXT_COL = 1, GXT_CONST = 2, GXT_K_SUB_COL = 3, GXT_K_ADD_COL = 4, GXT_K_MUL_COL = 5, GXT_COL_SUB_K = 6 };
struct gx_dterm { int op; int kind; gx_dcol col; double k; };
struct gx_dexpr { int nterms; int _pad; gx_dterm t[4]; };

enum { GXU_NONE = 0, GXU_ADD_F64 = 1, GXU_ADD_I64 = 2, GXU_MIN_F64 = 3, GXU_MAX_F64 = 4, GXU_CNT = 5 };
struct gx_dagg {
    int kind;          // GXU_*
    int word;          // state word receiving the value
    int cnt_word;      // word receiving +1 per non-NULL input (0 = none: w0 row count serves)
    int is_int;        // value is an integer column (SUM_I4/I8, COUNT(col))
    gx_dexpr expr;     // float8 argument (kind F64) ...
    gx_dcol icol;      // ... or integer argument column
};

struct gx_dgroupcol { int side; int type; gx_dcol col; int payload_idx; int word; int shift; int bytes; int _pad; };

struct gx_dplan {
    int npreds, nagg, ngroup, nkw;
    int nwords, has_join, key_type, unique;
    gx_dpred preds[GX_MAX_PREDS];
    gx_dcol okey;                       // outer join key column
    gx_dgroupcol gcols[GX_MAX_GROUP_COLS];
    gx_dagg aggs[GX_MAX_AGGS];
    int payload_types[GX_MAX_PAYLOAD];
    int n_payload; int _pad;
};

#ifdef __CUDACC__
// ---------------------------------------------------------------- device side

__device__ __forceinline__ unsigned int gx_rotl32(unsigned int x, int k) { return __funnelshift_l(x, x, k); }

// final() of Bob Jenkins' lookup3 as used by hash_uint32()
// (reference: src/backend/access/hash/hashfunc.c:593-602, 1044-1058)
__device__ __forceinline__ unsigned int gx_hash_uint32(unsigned int k)
{
    unsigned int a, b, c;
    a = b = c = 0x9e3779b9u + 4u + 3923095u;
    a += k;
    c ^= b; c -= gx_rotl32(b, 14);
    a ^= c; a -= gx_rotl32(c, 11);
    b ^= a; b -= gx_rotl32(a, 25);
    c ^= b; c -= gx_rotl32(b, 16);
    a ^= c; a -= gx_rotl32(c, 4);
    b ^= a; b -= gx_rotl32(a, 14);
    c ^= b; c -= gx_rotl32(b, 24);
    return c;
}
// hashint8(): fold hi into lo (hashfunc.c:92-110)
__device__ __forceinline__ unsigned int gx_hashint8(long long v)
{
    unsigned int lo = (unsigned int) v, hi = (unsigned int) ((unsigned long long) v >> 32);
    lo ^= (v >= 0) ? hi : ~hi;
    return gx_hash_uint32(lo);
}
__device__ __forceinline__ unsigned int gx_hashint4(int v) { return gx_hash_uint32((unsigned int) v); }

// CRC32C of the 8 little-endian bytes of v, bitwise (hash_any_new, hashfunc.c:112)
__device__ __forceinline__ unsigned int gx_crc32c_u64(unsigned long long v)
{
    unsigned int crc = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        crc ^= (unsigned int) (v >> (8 * i)) & 0xFFu;
#pragma unroll
        for (int j = 0; j < 8; j++) crc = (crc >> 1) ^ (0x82F63B78u & (0u - (crc & 1u)));
    }
    return crc ^ 0xFFFFFFFFu;
}
__device__ __forceinline__ unsigned int gx_murmurhash32(unsigned int h)
{
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}
// internal bucket hash of the join / group tables (need not match the reference:
// SURVEY.md §2 row 7): murmur3 fmix64
__device__ __forceinline__ unsigned long long gx_mix64(unsigned long long h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h;
}

// SHARD routing, bit-exact (locator.c:1611, shardmap.c:1147): one distribution column
__device__ __forceinline__ int gx_shard_index(unsigned int hashvalue)
{
    int h = (int) hashvalue;
    unsigned int mag = (h < 0) ? (0u - (unsigned int) h) : (unsigned int) h;
    int a = (int) mag;                       // abs(INT_MIN) stays INT_MIN
    return a % GX_SHARD_MAP_SHARD_NUM;       // INT_MIN % 4096 == 0
}
__device__ __forceinline__ unsigned int gx_route_hash(int type, long long datum, bool isnull)
{
    unsigned int hashkey = 0;                // rotl1(0) == 0
    if (!isnull) hashkey ^= (type == GX_INT8) ? gx_hashint8(datum) : gx_hashint4((int) datum);
    return hashkey;
}

// typed column access ------------------------------------------------------
__device__ __forceinline__ long long gx_load_int(const gx_dcol &c, long long r)
{
    switch (c.type) {
        case GX_INT4: case GX_DATE: return (long long) __ldg((const int *) c.data + r);
        case GX_INT8: return __ldg((const long long *) c.data + r);
        case GX_CHAR: return (long long) __ldg((const signed char *) c.data + r);
        default: return __ldg((const long long *) c.data + r);     // FLOAT8 bit pattern
    }
}
__device__ __forceinline__ double gx_load_f64(const gx_dcol &c, long long r)
{
    switch (c.type) {
        case GX_FLOAT8: return __ldg((const double *) c.data + r);
        case GX_INT4: case GX_DATE: return (double) __ldg((const int *) c.data + r);      // i4tod
        case GX_INT8: return (double) __ldg((const long long *) c.data + r);              // i8tod
        default: return (double) __ldg((const signed char *) c.data + r);
    }
}
__device__ __forceinline__ bool gx_is_null(const gx_dcol &c, long long r)
{
    return c.nulls != nullptr && __ldg(c.nulls + r) != 0;
}

// float8_cmp_internal: NaN sorts above everything and equals itself (float.c:1160)
__device__ __forceinline__ int gx_f8cmp(double a, double b)
{
    if (isnan(a)) return isnan(b) ? 0 : 1;
    if (isnan(b)) return -1;
    return a > b ? 1 : (a < b ? -1 : 0);
}
// does "x op y" hold, given c = sign(x - y)?  Branch-free: three bits per operator (c = -1, 0, +1), looked up in one word -
// the switch this replaces compiled to a jump table inside every tile loop (ncu: 8 % of gx_k_runagg's instructions)
#define GX_OP_BITS(op, lt, eq, gt) ((unsigned int) ((lt) | ((eq) << 1) | ((gt) << 2)) << (3 * (op)))
__device__ __forceinline__ bool gx_op_holds(int op, int c)
{
    constexpr unsigned int T = GX_OP_BITS(GX_LT, 1, 0, 0) | GX_OP_BITS(GX_LE, 1, 1, 0) | GX_OP_BITS(GX_EQ, 0, 1, 0) |
                               GX_OP_BITS(GX_GE, 0, 1, 1) | GX_OP_BITS(GX_GT, 0, 0, 1) | GX_OP_BITS(GX_NE, 1, 0, 1);
    return (T >> (3 * op + c + 1)) & 1u;
}
__device__ __forceinline__ bool gx_eval_pred(const gx_dpred &p, long long r)
{
    if (gx_is_null(p.col, r)) return false;          // strict operator: NULL fails the qual
    int c;
    if (p.col.type == GX_FLOAT8) c = gx_f8cmp(gx_load_f64(p.col, r), p.fval);
    else if (p.col.type == GX_CHAR) {                // charlt etc. compare as uint8
        unsigned char x = (unsigned char) gx_load_int(p.col, r), y = (unsigned char) p.ival;
        c = x > y ? 1 : (x < y ? -1 : 0);
    } else { long long x = gx_load_int(p.col, r); c = x > p.ival ? 1 : (x < p.ival ? -1 : 0); }
    return gx_op_holds(p.op, c);
}

// quals of a tile of K rows per lane: the predicate descriptor is decoded once, the K column loads are independent
template <int K>
__device__ __forceinline__ void pred_tile(const gx_dpred &p, const long long (&r)[K], bool (&ok)[K])
{
    if (p.col.nulls == nullptr && (p.col.type == GX_INT4 || p.col.type == GX_DATE)) {
        const int *c = (const int *) p.col.data; int x[K];
#pragma unroll
        for (int j = 0; j < K; j++) x[j] = ok[j] ? __ldg(c + r[j]) : 0;
#pragma unroll
        for (int j = 0; j < K; j++) ok[j] = ok[j] && gx_op_holds(p.op, (long long) x[j] > p.ival ? 1 : ((long long) x[j] < p.ival ? -1 : 0));
    } else if (p.col.nulls == nullptr && p.col.type == GX_FLOAT8) {
        const double *c = (const double *) p.col.data; double x[K];
#pragma unroll
        for (int j = 0; j < K; j++) x[j] = ok[j] ? __ldg(c + r[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < K; j++) ok[j] = ok[j] && gx_op_holds(p.op, gx_f8cmp(x[j], p.fval));
    } else {
#pragma unroll
        for (int j = 0; j < K; j++) if (ok[j]) ok[j] = gx_eval_pred(p, r[j]);
    }
}

// IEEE fp64 ops that ptxas must never contract into FMA: the reference computes
// float8mul/float8pl/float8mi as separate correctly-rounded operations.
__device__ __forceinline__ double gx_apply(int op, double a, double b)
{
    return op == GX_OP_ADD ? __dadd_rn(a, b) : (op == GX_OP_SUB ? __dsub_rn(a, b) : __dmul_rn(a, b));
}
__device__ __forceinline__ double gx_eval_term(const gx_dterm &t, long long r, bool &isnull)
{
    switch (t.kind) {
        case GXT_CONST: return t.k;
        case GXT_COL: isnull |= gx_is_null(t.col, r); return gx_load_f64(t.col, r);
        case GXT_K_SUB_COL: isnull |= gx_is_null(t.col, r); return __dsub_rn(t.k, gx_load_f64(t.col, r));
        case GXT_K_ADD_COL: isnull |= gx_is_null(t.col, r); return __dadd_rn(t.k, gx_load_f64(t.col, r));
        case GXT_K_MUL_COL: isnull |= gx_is_null(t.col, r); return __dmul_rn(t.k, gx_load_f64(t.col, r));
        default: isnull |= gx_is_null(t.col, r); return __dsub_rn(gx_load_f64(t.col, r), t.k);
    }
}
__device__ __forceinline__ double gx_eval_expr(const gx_dexpr &e, long long r, bool &isnull)
{
    double acc = gx_eval_term(e.t[0], r, isnull);
#pragma unroll
    for (int i = 1; i < 4; i++)
        if (i < e.nterms) acc = gx_apply(e.t[i].op, acc, gx_eval_term(e.t[i], r, isnull));
    return acc;
}

// join-table probe -------------------------------------------------------
__device__ __forceinline__ unsigned long long gx_key_hash(long long key) { return gx_mix64((unsigned long long) key); }
// Slot function.  A probe that misses L2 costs a whole 128-byte line of DRAM traffic
// (profiles/r01_ncu_*: ~4 sectors per distinct key), i.e. eight 16-byte slots, so it pays
// to keep keys that are probed together in the same line.
//   mode 0: slot = fmix64(key) & mask                      (any keys)
//   mode 1: slot = floor((key - kmin) * scale / 2^64)       order-preserving interpolation:
//           keys spread near-uniformly over [kmin, kmax] (serial primary keys; TPC-H order
//           keys) land in key order, so a probe side clustered on the key walks the table
//           almost sequentially.  Picked from a key-density sample at build time and kept
//           only if the chains measured while filling stay short (else rebuilt with mode 0).
// mode 1 maps keys to slots in key order (probes of a key-ordered outer side then walk
// the table front to back); `win` (0 or 2^w - 1) additionally scatters the low w slot bits
// with a multiplicative hash, which breaks up the pile-ups that locally bunched keys
// (TPC-H order keys: 8 used of every 32) cause under pure interpolation while keeping
// every key inside the same few cache lines and the same sub-table.
struct gx_slotfn { int mode; unsigned int win; long long kmin; unsigned long long scale; unsigned long long mask; unsigned int shift; unsigned int amask; };
__device__ __forceinline__ unsigned long long gx_slot_index(long long key, const gx_slotfn &f)
{
    // Home slots are aligned to amask + 1 slots (2 x 16 B or 4 x 8 B = one 32-byte sector), which a
    // prober can fetch with a single 256-bit load; linear probing is unchanged.
    if (f.mode == 0) return gx_mix64((unsigned long long) key) & f.mask & ~(unsigned long long) f.amask;
    if (f.mode == 2) {
        // key range below 2^32: the same interpolation in 32-bit arithmetic (one wide multiply)
        const unsigned int d = (unsigned int) ((unsigned long long) key - (unsigned long long) f.kmin);
        const unsigned long long s = (((unsigned long long) d * (unsigned int) f.scale) >> f.shift) & f.mask;
        return (s ^ (unsigned long long) (((d * 0x9E3779B9u) >> 24) & f.win)) & ~(unsigned long long) f.amask;
    }
    unsigned long long s = __umul64hi((unsigned long long) key - (unsigned long long) f.kmin, f.scale) & f.mask;
    return (s ^ ((unsigned long long) (((unsigned long long) key * 0x9E3779B97F4A7C15ULL) >> 40) & f.win)) & ~(unsigned long long) f.amask;
}
// next aligned pair of a probe sequence (same wrapping rule as gx_next_slot)
__device__ __forceinline__ unsigned long long gx_next_pair(unsigned long long s, unsigned long long mask)
{
    const unsigned long long w = mask < (GX_SUB - 1) ? mask : (unsigned long long) (GX_SUB - 1);
    return (s & ~w) | ((s + 2) & w);
}
__device__ __forceinline__ unsigned long long gx_next_quad(unsigned long long s, unsigned long long mask)
{
    const unsigned long long w = mask < (GX_SUB - 1) ? mask : (unsigned long long) (GX_SUB - 1);
    return (s & ~w) | ((s + 4) & w);
}
// next slot of a probe sequence: wraps inside the slot's sub-table (or the whole table when it is smaller)
__device__ __forceinline__ unsigned long long gx_next_slot(unsigned long long s, unsigned long long mask)
{
    const unsigned long long w = mask < (GX_SUB - 1) ? mask : (unsigned long long) (GX_SUB - 1);
    return (s & ~w) | ((s + 1) & w);
}

// Copy K selected rows of one column (row r[j] -> position dst[j]): the K loads are independent and issued together,
// the type is decoded once per column instead of once per row (projection / compaction / scatter kernels).
template <int K>
__device__ __forceinline__ void gx_copy_rows(const gx_dcol &in, void *out, uint8_t *out_nulls, const long long (&r)[K], const long long (&dst)[K], const bool (&keep)[K])
{
    if (in.type == GX_INT4 || in.type == GX_DATE) {
        int v[K];
#pragma unroll
        for (int j = 0; j < K; j++) v[j] = keep[j] ? __ldg((const int *) in.data + r[j]) : 0;
#pragma unroll
        for (int j = 0; j < K; j++) if (keep[j]) ((int *) out)[dst[j]] = v[j];
    } else if (in.type == GX_CHAR) {
        signed char v[K];
#pragma unroll
        for (int j = 0; j < K; j++) v[j] = keep[j] ? __ldg((const signed char *) in.data + r[j]) : 0;
#pragma unroll
        for (int j = 0; j < K; j++) if (keep[j]) ((signed char *) out)[dst[j]] = v[j];
    } else {
        long long v[K];
#pragma unroll
        for (int j = 0; j < K; j++) v[j] = keep[j] ? __ldg((const long long *) in.data + r[j]) : 0;
#pragma unroll
        for (int j = 0; j < K; j++) if (keep[j]) ((long long *) out)[dst[j]] = v[j];
    }
    if (out_nulls) {
#pragma unroll
        for (int j = 0; j < K; j++) if (keep[j]) out_nulls[dst[j]] = in.nulls ? in.nulls[r[j]] : 0;
    }
}

// block-wide exclusive scan of one value per thread (blockDim.x <= 1024)
__device__ __forceinline__ long long gx_block_exscan(long long v, long long *total, long long *smem /* 33 */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) smem[wid] = x;
    __syncthreads();
    if (wid == 0) {
        long long w = (lane < (int) ((blockDim.x + 31) >> 5)) ? smem[lane] : 0, z = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { long long y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
        smem[lane] = z - w;
        if (lane == 31) smem[32] = z;
    }
    __syncthreads();
    long long res = smem[wid] + x - v;
    *total = smem[32];
    __syncthreads();
    return res;
}
#endif  // __CUDACC__
