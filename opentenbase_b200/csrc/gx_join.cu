// gx_join.cu — K2 open-addressing hash build and K3 probe (materialising).
//
// Table: power-of-two array of 16-byte slots {int64 key, uint64 payload},
// linear probing, load factor in (0.33, 0.67].  Empty = key INT64_MIN; rows
// whose key IS INT64_MIN live in a small side list.  The payload holds up to
// eight bytes of carried build-side columns (or the build row number), so a
// probe hit costs one 32-byte sector and no second gather.
//
// Replaces MultiExecPrivateHash / ExecHashTableInsert (nodeHash.c:157,1828:
// chained buckets of MinimalTuple copies in 32 KB chunks) and
// ExecScanHashBucket / ExecHashJoinImpl (nodeHash.c:2174, nodeHashjoin.c:186).
#include "gx_internal.cuh"
#include <type_traits>

int gx_fill_dpreds(gx_ctx *ctx, const gx_table *t, int n_preds, const gx_pred *preds, gx_dpred *out);

struct gx_build_args {
    gx_dcol key;
    int npreds, n_payload;
    gx_dpred preds[GX_MAX_PREDS];
    gx_dcol payload[GX_MAX_PAYLOAD];
    long long nrows;
    gx_slot *slots; unsigned long long mask;
    unsigned long long *special; int special_cap;
    gx_slotfn sf;                 // gx_slot_index
    long long *counters;          // [0] entries inserted, [1] special count
};

__device__ __forceinline__ unsigned long long pack_payload(const gx_build_args &a, long long r)
{
    if (a.n_payload == 0) return (unsigned long long) r;
    unsigned long long p = 0; int shift = 0;
#pragma unroll
    for (int i = 0; i < GX_MAX_PAYLOAD; i++) {
        if (i >= a.n_payload) break;
        int t = a.payload[i].type;
        unsigned long long v = (unsigned long long) gx_load_int(a.payload[i], r);
        int bytes = (t == GX_INT8 || t == GX_FLOAT8) ? 8 : (t == GX_CHAR ? 1 : 4);
        if (bytes < 8) v &= (1ULL << (8 * bytes)) - 1;
        p |= v << shift;
        shift += 8 * bytes;
    }
    return p;
}

__global__ void __launch_bounds__(256) gx_k_hash_build(gx_build_args a)
{
    long long stride = (long long) gridDim.x * blockDim.x;
    long long inserted = 0;
    for (long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x; r < a.nrows; r += stride) {
        bool ok = !gx_is_null(a.key, r);               // hashStrict: NULL keys never match
#pragma unroll
        for (int p = 0; p < GX_MAX_PREDS; p++) if (p < a.npreds) ok = ok && gx_eval_pred(a.preds[p], r);
        if (!ok) continue;
        long long key = gx_load_int(a.key, r);
        unsigned long long payload = pack_payload(a, r);
        if (key == GX_EMPTY_KEY) {
            int idx = (int) atomicAdd((unsigned long long *) &a.counters[1], 1ULL);
            if (idx < a.special_cap) a.special[idx] = payload;
            continue;
        }
        unsigned long long s = gx_slot_index(key, a.sf);
        for (;;) {
            long long old = (long long) atomicCAS((unsigned long long *) &a.slots[s].key,
                                                  (unsigned long long) GX_EMPTY_KEY, (unsigned long long) key);
            if (old == GX_EMPTY_KEY) { a.slots[s].payload = payload; break; }
            s = gx_next_slot(s, a.mask);
        }
        inserted++;
    }
    // one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) inserted += __shfl_down_sync(0xffffffffu, inserted, o);
    if ((threadIdx.x & 31) == 0 && inserted) atomicAdd((unsigned long long *) &a.counters[0], (unsigned long long) inserted);
}

__global__ void gx_k_fill_slots(gx_slot *slots, long long n)
{
    long long stride = (long long) gridDim.x * blockDim.x;
    // 16-byte stores
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        longlong2 v; v.x = GX_EMPTY_KEY; v.y = 0;
        ((longlong2 *) slots)[i] = v;
    }
}


// ---------------------------------------------------------------------------
// Bucketed build.  Random CAS into a table much larger than L2 tops out at
// ~16-21 G inserts/s on B200 (profiles/r01_ubench_random_access.txt): every
// insert is a DRAM read-modify-write of a sector that was just cleared.  So the
// table is cut into sub-tables of GX_SUB slots (32 KB); linear probing wraps
// INSIDE a sub-table (all probe kernels use gx_next_slot()).  Build rows are
// bucketed by sub-table (sequential read, 16-byte scattered stores that
// write-combine in L2), then one CTA builds one sub-table entirely in shared
// memory and streams it out with coalesced 16-byte stores: the table is written
// exactly once, never read, never cleared, and sees no global atomics.
struct gx_bbuild_args {
    gx_build_args b;
    long long nsub;
    unsigned int *cursor;       // [nsub] rows scattered to each sub-table so far
    gx_slot *pairs;             // nsub fixed-capacity buckets of GX_SUB pairs (a fuller bucket cannot be built anyway)
    int *overflow;
    long long *start; int *unsorted;   // key-ordered build: first row of every sub-table
    gx_cslot *cslots;                  // compact output (gx_k_sorted_fill<.., true>)
};

__device__ __forceinline__ bool build_row_ok(const gx_build_args &a, long long r)
{
    bool ok = !gx_is_null(a.key, r);
#pragma unroll
    for (int p = 0; p < GX_MAX_PREDS; p++) if (p < a.npreds) ok = ok && gx_eval_pred(a.preds[p], r);
    return ok;
}

// Scatter (key, payload) pairs into their sub-table's bucket.  The position
// comes from one L2 atomic; the chain load -> atomic -> store is latency bound,
// so every thread keeps BSCAT rows in flight.
#define BSCAT 4
__global__ void __launch_bounds__(256) gx_k_bbuild_scatter(gx_bbuild_args a)
{
    const long long tile = (long long) blockDim.x * BSCAT;
    const long long ntiles = (a.b.nrows + tile - 1) / tile;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        long long key[BSCAT]; unsigned long long payload[BSCAT], sub[BSCAT]; unsigned int pos[BSCAT]; bool ok[BSCAT];
#pragma unroll
        for (int u = 0; u < BSCAT; u++) {
            long long r = t * tile + (long long) u * blockDim.x + threadIdx.x;     // coalesced per u
            ok[u] = r < a.b.nrows && build_row_ok(a.b, r);
            key[u] = ok[u] ? gx_load_int(a.b.key, r) : 0;
            payload[u] = ok[u] ? pack_payload(a.b, r) : 0;
        }
#pragma unroll
        for (int u = 0; u < BSCAT; u++) {
            if (ok[u] && key[u] == GX_EMPTY_KEY) {
                int idx = (int) atomicAdd((unsigned long long *) &a.b.counters[1], 1ULL);
                if (idx < a.b.special_cap) a.b.special[idx] = payload[u];
                ok[u] = false;
            }
            sub[u] = gx_slot_index(key[u], a.b.sf) >> GX_SUB_LOG2;
            if (ok[u]) pos[u] = atomicAdd(&a.cursor[sub[u]], 1u);
        }
#pragma unroll
        for (int u = 0; u < BSCAT; u++) {
            if (!ok[u]) continue;
            if (pos[u] >= GX_SUB) { *a.overflow = 1; continue; }
            longlong2 v; v.x = key[u]; v.y = (long long) payload[u];
            ((longlong2 *) a.pairs)[sub[u] * GX_SUB + pos[u]] = v;
        }
    }
}
// ---------------------------------------------------------------------------
// Two-level write-combined bucketing (replaces the one-pass scatter above for
// big tables).  Measured: the one-pass scatter spends ~2.1 ms in per-row L2
// atomics and ~2.1 ms in 16-byte scattered stores at SF100.  Here a CTA takes a
// tile of 4096 rows, ranks them by bucket digit in shared memory (native 32-bit
// shared atomics), stages them bucket by bucket and copies each bucket's run out
// as one contiguous piece: one global atomic per (tile, bucket) instead of one
// per row, and 256-byte runs instead of 16-byte stores.  Level 1 splits on the
// high bits of the sub-table number, level 2 on the low bits.
#define BP_TILE 4096
#define BP_THREADS 512
#define BP_PER (BP_TILE / BP_THREADS)
struct gx_bpart_args {
    gx_build_args b;
    int bits_hi, bits_lo;
    long long nsub;
    gx_slot *l1; long long cap1; unsigned int *cur1;      // level-1 regions
    gx_slot *pairs; unsigned int *cursor; int *overflow;  // final buckets
    long long tiles_per_l1;
};

struct bp_smem {
    longlong2 *stage; unsigned short *bid; unsigned int *hist, *base, *gbase;
};
__device__ __forceinline__ bp_smem bp_carve(unsigned char *raw, int nb)
{
    bp_smem m;
    m.stage = (longlong2 *) raw;
    m.hist = (unsigned int *) (raw + (size_t) BP_TILE * 16);
    m.base = m.hist + nb; m.gbase = m.base + nb;
    m.bid = (unsigned short *) (m.gbase + nb);
    return m;
}
static size_t bp_smem_bytes(int nb) { return (size_t) BP_TILE * 16 + (size_t) nb * 12 + (size_t) BP_TILE * 2 + 64; }

// rank + stage one tile; returns after the staging buffer is complete
__device__ __forceinline__ void bp_stage_tile(const bp_smem &m, int nb, const long long *key, const unsigned long long *payload,
                                              const bool *ok, const unsigned int *digit, long long *scan_smem)
{
    unsigned int rank[BP_PER];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) m.hist[i] = 0;
    __syncthreads();
    // Sorted build input (serial keys + order-preserving slots) sends a whole warp to the same
    // bucket: aggregate per warp first so the shared atomic is issued once per distinct bucket.
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < BP_PER; u++) {
        unsigned int am = __ballot_sync(0xffffffffu, ok[u]);
        if (!ok[u]) continue;
        unsigned int peers = __match_any_sync(am, digit[u]);
        int leader = __ffs(peers) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&m.hist[digit[u]], (unsigned int) __popc(peers));
        rank[u] = __shfl_sync(peers, base, leader) + __popc(peers & ((1u << lane) - 1));
    }
    __syncthreads();
    {   // exclusive scan of the histogram (nb <= blockDim.x)
        long long tot;
        long long h = (int) threadIdx.x < nb ? (long long) m.hist[threadIdx.x] : 0;
        long long ex = gx_block_exscan(h, &tot, scan_smem);
        if ((int) threadIdx.x < nb) m.base[threadIdx.x] = (unsigned int) ex;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < BP_PER; u++) {
        if (!ok[u]) continue;
        unsigned int p = m.base[digit[u]] + rank[u];
        longlong2 v; v.x = key[u]; v.y = (long long) payload[u];
        m.stage[p] = v; m.bid[p] = (unsigned short) digit[u];
    }
}

__global__ void __launch_bounds__(BP_THREADS, 2) gx_k_bpart1(gx_bpart_args a)
{
    extern __shared__ unsigned char bp_raw[];
    __shared__ long long scan_smem[33];
    const int nb = 1 << a.bits_hi;
    bp_smem m = bp_carve(bp_raw, nb);
    const long long ntiles = (a.b.nrows + BP_TILE - 1) / BP_TILE;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        long long key[BP_PER]; unsigned long long payload[BP_PER]; bool ok[BP_PER]; unsigned int digit[BP_PER];
#pragma unroll
        for (int u = 0; u < BP_PER; u++) {
            long long r = t * BP_TILE + (long long) u * BP_THREADS + threadIdx.x;
            ok[u] = r < a.b.nrows && build_row_ok(a.b, r);
            key[u] = ok[u] ? gx_load_int(a.b.key, r) : 0;
            payload[u] = ok[u] ? pack_payload(a.b, r) : 0;
            if (ok[u] && key[u] == GX_EMPTY_KEY) {
                int idx = (int) atomicAdd((unsigned long long *) &a.b.counters[1], 1ULL);
                if (idx < a.b.special_cap) a.b.special[idx] = payload[u];
                ok[u] = false;
            }
            digit[u] = (unsigned int) ((gx_slot_index(key[u], a.b.sf) >> GX_SUB_LOG2) >> a.bits_lo);
        }
        bp_stage_tile(m, nb, key, payload, ok, digit, scan_smem);
        if ((int) threadIdx.x < nb) m.gbase[threadIdx.x] = m.hist[threadIdx.x] ? atomicAdd(&a.cur1[threadIdx.x], m.hist[threadIdx.x]) : 0;
        __syncthreads();
        const unsigned int total = m.base[nb - 1] + m.hist[nb - 1];
        for (unsigned int j = threadIdx.x; j < total; j += blockDim.x) {
            unsigned int d = m.bid[j];
            long long pos = (long long) m.gbase[d] + (j - m.base[d]);
            if (pos >= a.cap1) { *a.overflow = 1; continue; }
            ((longlong2 *) a.l1)[(long long) d * a.cap1 + pos] = m.stage[j];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BP_THREADS, 2) gx_k_bpart2(gx_bpart_args a)
{
    extern __shared__ unsigned char bp_raw[];
    __shared__ long long scan_smem[33];
    const int nb = 1 << a.bits_lo;
    bp_smem m = bp_carve(bp_raw, nb);
    const long long nwork = ((long long) 1 << a.bits_hi) * a.tiles_per_l1;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
        const long long b1 = w / a.tiles_per_l1, t = w % a.tiles_per_l1;
        long long n1 = a.cur1[b1]; if (n1 > a.cap1) n1 = a.cap1;
        if (t * BP_TILE >= n1) continue;                       // uniform per CTA
        const longlong2 *src = (const longlong2 *) a.l1 + b1 * a.cap1;
        long long key[BP_PER]; unsigned long long payload[BP_PER]; bool ok[BP_PER]; unsigned int digit[BP_PER];
#pragma unroll
        for (int u = 0; u < BP_PER; u++) {
            long long i = t * BP_TILE + (long long) u * BP_THREADS + threadIdx.x;
            ok[u] = i < n1;
            longlong2 v; v.x = 0; v.y = 0;
            if (ok[u]) v = src[i];
            key[u] = v.x; payload[u] = (unsigned long long) v.y;
            digit[u] = (unsigned int) ((gx_slot_index(key[u], a.b.sf) >> GX_SUB_LOG2) & (unsigned long long) (nb - 1));
        }
        bp_stage_tile(m, nb, key, payload, ok, digit, scan_smem);
        if ((int) threadIdx.x < nb)
            m.gbase[threadIdx.x] = m.hist[threadIdx.x] ? atomicAdd(&a.cursor[(b1 << a.bits_lo) | threadIdx.x], m.hist[threadIdx.x]) : 0;
        __syncthreads();
        const unsigned int total = m.base[nb - 1] + m.hist[nb - 1];
        for (unsigned int j = threadIdx.x; j < total; j += blockDim.x) {
            unsigned int d = m.bid[j];
            unsigned int pos = m.gbase[d] + (j - m.base[d]);
            if (pos >= GX_SUB) { *a.overflow = 1; continue; }
            ((longlong2 *) a.pairs)[(((long long) b1 << a.bits_lo) | d) * GX_SUB + pos] = m.stage[j];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Building one sub-table (GX_SUB slots) in shared memory WITHOUT a compare-and-swap loop.
// A linear-probing table filled in slot order has a closed form: with c[s] rows hashing to
// slot s, the first of them lands at s + e[s], where e[s] = max(0, e[s-1] + c[s-1] - 1) is
// the overflow carried in from the left.  With P[s] = sum_{t<s} (c[t] - 1) this is
// e[s] = P[s] - min_{j<=s} P[j]: one histogram (native 32-bit shared atomics, which also
// hand every row its rank among the rows of its slot), one block-wide scan of the pair
// (sum, prefix-minimum), one conflict-free placement pass.  Rows pushed past the end of
// the sub-table wrap to its start; they are rare and go through a CAS loop afterwards.
// Any prober that walks from slot(key) to the right (gx_next_slot) finds every key before
// it meets an empty slot, exactly as with CAS insertion.
#define FILL_THREADS 256
#define FILL_ROWS (GX_SUB / FILL_THREADS)          // rows (and slots) per thread
template <bool C> struct gx_fill_smem_t {
    typename std::conditional<C, gx_cslot, gx_slot>::type tab[GX_SUB];
    unsigned int cnt[GX_SUB];                      // histogram, then e[s]
    int wsum[FILL_THREADS / 32], wmin[FILL_THREADS / 32];
};
typedef gx_fill_smem_t<false> gx_fill_smem;
#define FILL_SMEM_BYTES ((int) sizeof(gx_fill_smem))
#define FILL_SMEM_BYTES_C ((int) sizeof(gx_fill_smem_t<true>))
#define FILL_INF (1 << 29)

// C = compact 8-byte slots {key - kmin + 1, payload}; dst then points at gx_cslot
template <bool C, class LOAD>
__device__ __forceinline__ void gx_subtable_build(gx_fill_smem_t<C> &sm, unsigned int n, const gx_slotfn &sf, LOAD load, void *dst,
                                                  unsigned int &steps, unsigned int &placed, long long want_sub = -1, int *misplaced = nullptr)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- clear: histogram and table
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *c4 = (uint4 *) sm.cnt;
        for (int i = tid; i < GX_SUB / 4; i += FILL_THREADS) c4[i] = z;
        if (C) {
            uint4 *t4 = (uint4 *) sm.tab;
#pragma unroll
            for (int u = 0; u < FILL_ROWS / 2; u++) t4[tid + u * FILL_THREADS] = z;
        } else {
            longlong2 e; e.x = GX_EMPTY_KEY; e.y = 0;
            longlong2 *t2 = (longlong2 *) sm.tab;
#pragma unroll
            for (int u = 0; u < FILL_ROWS; u++) t2[tid + u * FILL_THREADS] = e;
        }
    }
    __syncthreads();
    // ---- histogram + rank
    // what is kept per row until placement: the slot's two words (compact: two 32-bit values)
    typename std::conditional<C, unsigned int, long long>::type key[FILL_ROWS];
    typename std::conditional<C, unsigned int, unsigned long long>::type pay[FILL_ROWS];
    unsigned int sr[FILL_ROWS];
#pragma unroll
    for (int u = 0; u < FILL_ROWS; u++) {
        unsigned int i = tid + u * FILL_THREADS;
        sr[u] = 0xffffffffu;
        long long k64; unsigned long long p64;
        if (i < n && load(i, k64, p64)) {
            const unsigned long long fs = gx_slot_index(k64, sf);
            if (want_sub >= 0 && (long long) (fs >> GX_SUB_LOG2) != want_sub) *misplaced = 1;   // caller's row range was wrong
            unsigned int sl = (unsigned int) (fs & (GX_SUB - 1));
            sr[u] = sl | (atomicAdd(&sm.cnt[sl], 1u) << GX_SUB_LOG2);
            if (C) { key[u] = GX_CSLOT_D(k64, sf.kmin); pay[u] = (unsigned int) p64; }
            else { key[u] = k64; pay[u] = p64; }
        }
    }
    __syncthreads();
    // ---- e[s] for the FILL_ROWS slots this thread owns
    {
        int c[FILL_ROWS];
        const uint4 *c4 = (const uint4 *) (sm.cnt + tid * FILL_ROWS);
#pragma unroll
        for (int q = 0; q < FILL_ROWS / 4; q++) { uint4 v = c4[q]; c[4 * q] = (int) v.x; c[4 * q + 1] = (int) v.y; c[4 * q + 2] = (int) v.z; c[4 * q + 3] = (int) v.w; }
        int p[FILL_ROWS], run = 0, lmin = 0;
#pragma unroll
        for (int j = 0; j < FILL_ROWS; j++) { p[j] = run; lmin = min(lmin, run); run += c[j] - 1; }
        // inclusive warp scan of (sum, prefix-min) with combine(a, b) = (a.s + b.s, min(a.m, a.s + b.m))
        int S = run, M = lmin;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int ps = __shfl_up_sync(0xffffffffu, S, o), pm = __shfl_up_sync(0xffffffffu, M, o);
            if (lane >= o) { M = min(pm, ps + M); S = ps + S; }
        }
        if (lane == 31) { sm.wsum[warp] = S; sm.wmin[warp] = M; }
        int xs = __shfl_up_sync(0xffffffffu, S, 1), xm = __shfl_up_sync(0xffffffffu, M, 1);
        if (lane == 0) { xs = 0; xm = FILL_INF; }
        __syncthreads();
        int bs = 0, bm = FILL_INF;                    // everything in the warps before this one
        for (int w = 0; w < warp; w++) { bm = min(bm, bs + sm.wmin[w]); bs += sm.wsum[w]; }
        const int base = bs + xs, mprev = min(bm, bs + xm);   // P at this thread's first slot; min of P over all earlier slots
        int rmin = mprev;
        uint4 o4[FILL_ROWS / 4];
        unsigned int *o = (unsigned int *) o4;
#pragma unroll
        for (int j = 0; j < FILL_ROWS; j++) { int P = base + p[j]; rmin = min(rmin, P); o[j] = (unsigned int) (P - rmin); }
        uint4 *d4 = (uint4 *) (sm.cnt + tid * FILL_ROWS);
#pragma unroll
        for (int q = 0; q < FILL_ROWS / 4; q++) d4[q] = o4[q];
    }
    __syncthreads();
    // ---- placement
    int wrapped = 0;
#pragma unroll
    for (int u = 0; u < FILL_ROWS; u++) {
        if (sr[u] == 0xffffffffu) continue;
        unsigned int sl = sr[u] & (GX_SUB - 1), pos = sl + sm.cnt[sl] + (sr[u] >> GX_SUB_LOG2);
        steps += pos - sl; placed++;
        if (pos < GX_SUB) {
            if (C) { uint2 v; v.x = (unsigned int) key[u]; v.y = (unsigned int) pay[u]; ((uint2 *) sm.tab)[pos] = v; }
            else { longlong2 v; v.x = key[u]; v.y = (long long) pay[u]; ((longlong2 *) sm.tab)[pos] = v; }
            sr[u] = 0xffffffffu;
        } else wrapped = 1;
    }
    if (__syncthreads_or(wrapped)) {
#pragma unroll
        for (int u = 0; u < FILL_ROWS; u++) {
            if (sr[u] == 0xffffffffu) continue;
            unsigned int sl = 0;                      // everything from its slot to the end is full: continue from the start
            for (;;) {
                bool won;
                if (C) {
                    gx_cslot *t = (gx_cslot *) sm.tab;
                    won = atomicCAS(&t[sl].d, 0u, (unsigned int) key[u]) == 0u;
                    if (won) t[sl].payload = (unsigned int) pay[u];
                } else {
                    gx_slot *t = (gx_slot *) sm.tab;
                    won = (long long) atomicCAS((unsigned long long *) &t[sl].key, (unsigned long long) GX_EMPTY_KEY, (unsigned long long) key[u]) == GX_EMPTY_KEY;
                    if (won) t[sl].payload = pay[u];
                }
                if (won) break;
                sl = (sl + 1) & (GX_SUB - 1); steps++;
            }
        }
        __syncthreads();
    }
    // ---- stream the finished sub-table out
    if (C) {
        const uint4 *t4 = (const uint4 *) sm.tab;
        uint4 *d4 = (uint4 *) dst;
#pragma unroll
        for (int u = 0; u < FILL_ROWS / 2; u++) d4[tid + u * FILL_THREADS] = t4[tid + u * FILL_THREADS];
    } else {
        const longlong2 *t2 = (const longlong2 *) sm.tab;
        longlong2 *d2 = (longlong2 *) dst;
#pragma unroll
        for (int u = 0; u < FILL_ROWS; u++) d2[tid + u * FILL_THREADS] = t2[tid + u * FILL_THREADS];
    }
    __syncthreads();
}

__device__ __forceinline__ void fill_report(const gx_build_args &b, unsigned int steps, unsigned int placed)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { steps += __shfl_down_sync(0xffffffffu, steps, o); placed += __shfl_down_sync(0xffffffffu, placed, o); }
    if ((threadIdx.x & 31) == 0) {
        if (steps) atomicAdd((unsigned long long *) &b.counters[5], (unsigned long long) steps);      // chain statistics decide whether
        if (placed) atomicAdd((unsigned long long *) &b.counters[3], (unsigned long long) placed);    // the slot function is kept
    }
}

// one CTA per sub-table (grid-strided) over the buckets the scatter passes produced
__global__ void __launch_bounds__(FILL_THREADS, 4) gx_k_bbuild_fill(gx_bbuild_args a)
{
    extern __shared__ __align__(16) unsigned char fill_smem_raw[];
    gx_fill_smem &sm = *(gx_fill_smem *) fill_smem_raw;
    unsigned int steps = 0, placed = 0;
    for (long long sub = blockIdx.x; sub < a.nsub; sub += gridDim.x) {
        unsigned int n = a.cursor[sub];
        if (n > GX_SUB) n = 0;                         // overflowed: the host rebuilds directly
        const longlong2 *src = (const longlong2 *) a.pairs + sub * GX_SUB;
        gx_subtable_build<false>(sm, n, a.b.sf,
                          [&](unsigned int i, long long &k, unsigned long long &p) { longlong2 v = src[i]; k = v.x; p = (unsigned long long) v.y; return true; },
                          a.b.slots + sub * GX_SUB, steps, placed);
    }
    fill_report(a.b, steps, 0);                        // entries are counted from the cursors (gx_k_bbuild_total)
}

// ---------------------------------------------------------------------------
// Partition-free build for a build side that is stored in key order (a table kept in
// primary-key order, e.g. TPC-H orders) when the order-preserving slot function is in
// use: the rows of one sub-table are then one CONTIGUOUS row range, found by binary
// search, so no bucketing pass is needed at all — each CTA reads its rows straight from
// the columns (coalesced), builds the sub-table in shared memory and streams it out.
// one coalesced pass over the key column: verifies ascending order and records, for every
// sub-table, the first row that maps to it (start[nsub] = nrows)
__global__ void gx_k_sorted_bounds(gx_build_args a, long long nsub, long long *start, int *unsorted)
{
    long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < a.nrows; i += stride) {
        if (gx_is_null(a.key, i)) { *unsorted = 1; return; }
        long long k = gx_load_int(a.key, i);
        long long sb = (k == GX_EMPTY_KEY) ? -1 : (long long) (gx_slot_index(k, a.sf) >> GX_SUB_LOG2);
        long long sp = -1;
        if (i > 0) {
            long long kp = gx_load_int(a.key, i - 1);
            if (kp > k) { *unsorted = 1; return; }
            sp = (kp == GX_EMPTY_KEY) ? -1 : (long long) (gx_slot_index(kp, a.sf) >> GX_SUB_LOG2);
        }
        for (long long s2 = sp + 1; s2 <= sb; s2++) start[s2] = i;
        if (i == a.nrows - 1) for (long long s2 = sb + 1; s2 <= nsub; s2++) start[s2] = a.nrows;
    }
}

// the same for an int8 key column without NULLs: 4 consecutive keys per lane (two 16-byte
// loads), the predecessor of a lane's first key comes from the lane below by shuffle
__global__ void __launch_bounds__(256) gx_k_sorted_bounds_i8(const long long *__restrict__ keys, long long nrows, gx_slotfn sf,
                                                             long long nsub, long long *start, int *unsorted)
{
    const int lane = threadIdx.x & 31;
    const long long nwarp = ((long long) gridDim.x * blockDim.x) >> 5, wid = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long ngroups = (nrows + 127) >> 7;
    bool bad = false;
    for (long long gq = wid; gq < ngroups; gq += nwarp) {
        const long long r0 = (gq << 7) + lane * 4;
        long long k[4];
        if (r0 + 3 < nrows) {
            longlong2 a = __ldg((const longlong2 *) (keys + r0)), b = __ldg((const longlong2 *) (keys + r0 + 2));
            k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) k[j] = r0 + j < nrows ? __ldg(keys + r0 + j) : 0x7fffffffffffffffLL;
        }
        long long sb[4];
#pragma unroll
        for (int j = 0; j < 4; j++) sb[j] = (k[j] == GX_EMPTY_KEY) ? -1 : (long long) (gx_slot_index(k[j], sf) >> GX_SUB_LOG2);
        long long kp = __shfl_up_sync(0xffffffffu, k[3], 1), sp = __shfl_up_sync(0xffffffffu, sb[3], 1);
        if (lane == 0) {
            if (r0 == 0) { kp = k[0]; sp = -1; }
            else { kp = __ldg(keys + r0 - 1); sp = (kp == GX_EMPTY_KEY) ? -1 : (long long) (gx_slot_index(kp, sf) >> GX_SUB_LOG2); }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long r = r0 + j;
            if (r < nrows) {
                if (kp > k[j]) bad = true;
                if (sb[j] != sp) for (long long s2 = sp + 1; s2 <= sb[j]; s2++) start[s2] = r;
                if (r == nrows - 1) for (long long s2 = sb[j] + 1; s2 <= nsub; s2++) start[s2] = nrows;
                kp = k[j]; sp = sb[j];
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) *unsorted = 1;
}

// PK selects the row loader: 0 generic (any key type, NULLs, build-side quals, packed payload),
// 1 = int8 key without NULLs or quals + one 4-byte payload column, 2 = the same without payload
template <int PK, bool C, bool VERIFY>
__global__ void __launch_bounds__(FILL_THREADS, C ? 5 : 4) gx_k_sorted_fill(gx_bbuild_args a)      // 5 CTAs/SM: profiles/r02_occupancy_variants.txt
{
    extern __shared__ __align__(16) unsigned char fill_smem_raw[];
    gx_fill_smem_t<C> &sm = *(gx_fill_smem_t<C> *) fill_smem_raw;
    void *const out = C ? (void *) a.cslots : (void *) a.b.slots;
    const size_t slot_bytes = C ? sizeof(gx_cslot) : sizeof(gx_slot);
    if (!VERIFY && *a.unsorted) return;
    unsigned int steps = 0, placed = 0;
    int bad = 0;
    const long long *keys = (const long long *) a.b.key.data;
    const unsigned int *pay4 = (const unsigned int *) a.b.payload[0].data;
    long long lo = 0, hi = 0;
    if (blockIdx.x < a.nsub) { lo = a.start[blockIdx.x]; hi = a.start[blockIdx.x + 1]; }
    for (long long sub = blockIdx.x; sub < a.nsub; sub += gridDim.x) {
        // the next sub-table's row range: fetch its bounds now and pull its rows towards L2
        // while this one is being built
        long long nlo = 0, nhi = 0;
        const long long nsubn = sub + gridDim.x;
        if (nsubn < a.nsub) { nlo = a.start[nsubn]; nhi = a.start[nsubn + 1]; }
        unsigned int n = (unsigned int) (hi - lo);
        if (hi - lo > GX_SUB) { n = 0; if (threadIdx.x == 0) *a.overflow = 1; }
        if (VERIFY && hi < lo) { n = 0; bad = 1; }
        if (PK != 0 && nhi - nlo <= GX_SUB) {
            const long long l = nlo + (long long) threadIdx.x * 16;          // 16 keys = one 128-byte line
            if (l < nhi) asm volatile("prefetch.global.L2 [%0];" :: "l"(keys + l));
            if (PK == 1) { const long long l4 = nlo + (long long) threadIdx.x * 32; if (l4 < nhi) asm volatile("prefetch.global.L2 [%0];" :: "l"(pay4 + l4)); }
        }
        if (PK == 0 && !VERIFY)
            gx_subtable_build<C>(sm, n, a.b.sf,
                              [&](unsigned int i, long long &k, unsigned long long &p) {
                                  long long r = lo + i;
                                  if (!build_row_ok(a.b, r)) return false;
                                  k = gx_load_int(a.b.key, r); p = pack_payload(a.b, r); return true; },
                              (char *) out + (size_t) sub * GX_SUB * slot_bytes, steps, placed);
        else if (PK == 0)                                      // int8 key without NULLs; quals and any payload
            gx_subtable_build<C>(sm, n, a.b.sf,
                              [&](unsigned int i, long long &k, unsigned long long &p) {
                                  const long long r = lo + i;
                                  k = __ldg(keys + r);
                                  const long long kp = r > 0 ? __ldg(keys + r - 1) : k;
                                  if (kp > k || k == GX_EMPTY_KEY) bad = 1;          // order is checked on every row, kept or not
                                  if (!build_row_ok(a.b, r)) return false;
                                  p = pack_payload(a.b, r); return true; },
                              (char *) out + (size_t) sub * GX_SUB * slot_bytes, steps, placed, sub, &bad);
        else if (!VERIFY)
            gx_subtable_build<C>(sm, n, a.b.sf,
                              [&](unsigned int i, long long &k, unsigned long long &p) {
                                  k = __ldg(keys + lo + i);
                                  p = PK == 1 ? (unsigned long long) __ldg(pay4 + lo + i) : (unsigned long long) (lo + i);
                                  return true; },
                              (char *) out + (size_t) sub * GX_SUB * slot_bytes, steps, placed);
        else
            gx_subtable_build<C>(sm, n, a.b.sf,
                              [&](unsigned int i, long long &k, unsigned long long &p) {
                                  const long long r = lo + i;
                                  k = __ldg(keys + r);
                                  const long long kp = r > 0 ? __ldg(keys + r - 1) : k;    // the neighbour's load: an L1 hit
                                  if (kp > k || k == GX_EMPTY_KEY) bad = 1;
                                  p = PK == 1 ? (unsigned long long) __ldg(pay4 + r) : (unsigned long long) r;
                                  return true; },
                              (char *) out + (size_t) sub * GX_SUB * slot_bytes, steps, placed, sub, &bad);
        lo = nlo; hi = nhi;
    }
    if (VERIFY && bad) *a.unsorted = 1;
    fill_report(a.b, steps, placed);
}
// ---------------------------------------------------------------------------
// Bounds by SEARCH instead of by a pass over the key column (int8 key without NULLs).
//  * lower_bound(s) = first row whose key maps to sub-table >= s.  One warp: the 32 rows around
//    the interpolated position s * rows-per-sub-table decide it in one round when the keys are
//    as uniform as the slot function assumes; otherwise a bracket is grown (x16 per round) and
//    narrowed 32-ary.  All nsub + 1 bounds cost microseconds (gx_k_sorted_bounds_search).
//  * what the pass over the column used to establish is then VERIFIED by the fill itself
//    (gx_k_sorted_fill<.., VERIFY>): every CTA checks that its rows map to its sub-table and
//    ascend (each row against its predecessor, including the one before its range).
//    start[0] = 0 and start[nsub] = nrows are fixed and the ranges [start[s], start[s+1]) tile
//    [0, nrows), so if no CTA raises the flag every adjacent pair of rows has been compared and
//    every row sits in the right sub-table — whatever the search returned.  A raised flag sends
//    the host down the bucketing path.  (Doing the search inside the fill kernel, one sub-table
//    ahead, was measured slower: 1.91 ms against 0.41 + 1.07 ms.)
__device__ __forceinline__ long long sorted_lower_bound(const long long *__restrict__ keys, long long nrows, const gx_slotfn &sf,
                                                        long long s, long long nsub, double rows_per_sub, int lane, long long seed = -1)
{
    if (s <= 0) return 0;
    if (s >= nsub) return nrows;
    auto pred = [&](long long r) -> bool {                      // true from the bound onwards
        if (r >= nrows) return true;
        if (r < 0) return false;
        return (long long) (gx_slot_index(__ldg(keys + r), sf) >> GX_SUB_LOG2) >= s;
    };
    // guess: the global density, or — better — the previous bound plus one sub-table's worth of rows (a datanode's 1/N
    // pseudo-random subset of the keys wanders off the global density by hundreds of rows, but only by a dozen per sub-table)
    long long g = seed >= 0 ? seed + (long long) rows_per_sub : (long long) ((double) s * rows_per_sub);
    if (g > nrows) g = nrows;
    // one round when the guess is within 16 rows
    {
        const long long pos = g - 16;
        const unsigned int b = __ballot_sync(0xffffffffu, pred(pos + lane));
        if (b != 0u && b != 0xffffffffu) return pos + (__ffs((int) b) - 1);
    }
    // grow a bracket [lo, hi]: pred(lo - 1) false (or lo == 0), pred(hi) true (or hi == nrows)
    long long lo = 0, hi = nrows;
    for (long long e = 256; ; e *= 16) {
        const long long cl = g - e > 0 ? g - e : 0, ch = g + e < nrows ? g + e : nrows;
        const bool p = lane == 0 ? pred(cl - 1) : (lane == 1 ? pred(ch) : false);
        const unsigned int b = __ballot_sync(0xffffffffu, p);
        const bool lo_ok = cl == 0 || !(b & 1u), hi_ok = ch == nrows || (b & 2u);
        if (lo_ok) lo = cl;
        if (hi_ok) hi = ch;
        if ((lo_ok && hi_ok) || (cl == 0 && ch == nrows)) break;
    }
    // 32-ary narrowing
    for (int round = 0; round < 16 && hi - lo > 32; round++) {
        const long long step = (hi - lo + 31) / 32;
        const long long r = lo + (long long) lane * step;
        const unsigned int b = __ballot_sync(0xffffffffu, r >= hi ? true : pred(r));
        if (b == 0u) { lo = lo + 31 * step + 1; continue; }
        const int f = __ffs((int) b) - 1;
        hi = lo + (long long) f * step < hi ? lo + (long long) f * step : hi;
        if (f > 0) lo = lo + (long long) (f - 1) * step + 1;
    }
    {
        const long long r = lo + lane;
        const unsigned int b = __ballot_sync(0xffffffffu, r >= hi ? true : pred(r));
        return b ? lo + (__ffs((int) b) - 1) : hi;
    }
}

// start[s] for every sub-table by search: one warp per bound, ~one 256-byte read each
__global__ void __launch_bounds__(256) gx_k_sorted_bounds_search(const long long *__restrict__ keys, long long nrows, gx_slotfn sf,
                                                                 long long nsub, long long nslots, long long *start)
{
    const int lane = threadIdx.x & 31;
    const long long nwarp = ((long long) gridDim.x * blockDim.x) >> 5, wid = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const double rows_per_sub = (double) nrows * (double) GX_SUB / (double) (nslots - GX_SUB);
    // a warp takes a CONTIGUOUS block of bounds and seeds each search with the previous result
    const long long per = (nsub + 1 + nwarp - 1) / nwarp, s0 = wid * per;
    long long prev = -1;
    for (long long s = s0; s < s0 + per && s <= nsub; s++) {
        const long long v = sorted_lower_bound(keys, nrows, sf, s, nsub, rows_per_sub, lane, prev);
        if (lane == 0) start[s] = v;
        prev = (s > 0 && s < nsub) ? v : -1;
    }
}

// rows with the reserved key INT64_MIN sort first: move them to the side list
__global__ void gx_k_sorted_special(gx_build_args a, const long long *start, const int *unsorted)
{
    if (*unsorted) return;
    const long long n = start[0];
    for (long long r = threadIdx.x; r < n; r += blockDim.x) {
        if (gx_is_null(a.key, r) || gx_load_int(a.key, r) != GX_EMPTY_KEY) continue;
        if (!build_row_ok(a, r)) continue;
        int idx = (int) atomicAdd((unsigned long long *) &a.counters[1], 1ULL);
        if (idx < a.special_cap) a.special[idx] = pack_payload(a, r);
    }
}

// entries actually placed = sum of the cursors
__global__ void gx_k_bbuild_total(const unsigned int *cursor, long long n, long long *total)
{
    long long sum = 0;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) sum += cursor[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd((unsigned long long *) total, (unsigned long long) sum);
}
__global__ void gx_k_scan_i64(long long *v, long long n, long long *total);   // gx_agg.cu

// key-density estimate for the slot function: min / max over a strided sample
__global__ void gx_k_key_range(gx_dcol key, long long nrows, long long stride, long long *minmax)
{
    long long i = ((long long) blockIdx.x * blockDim.x + threadIdx.x) * stride;
    long long lo = 0x7fffffffffffffffLL, hi = -0x7fffffffffffffffLL - 1;
    if (i < nrows && !gx_is_null(key, i)) { lo = hi = gx_load_int(key, i); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long a = __shfl_down_sync(0xffffffffu, lo, o), b = __shfl_down_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo; hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&minmax[0], lo); atomicMax(&minmax[1], hi); }
}

// Interpolation is tried when the key span is at most this many times the row count: TPC-H order
// keys use 8 of every 32 values (spread 4), and each of N datanodes holds a pseudo-random 1/N of
// them (spread 4 N).  Uniformity itself is not sampled: a build whose sub-tables overflow or whose
// average chain exceeds 4 is redone with the mixing hash.
#define GX_INTERP_MAX_SPREAD 256.0

// interpolation slot function over [kmin, kmin + range): slots [0, nslots - GX_SUB), monotone in the key.
// Narrow ranges (< 2^32, the usual case: at most 16 keys of range per row) use 32-bit arithmetic.
static void gx_set_interpolation(gx_hash *h, long long kmin, double range_d)
{
    const long double ratio = (long double) (h->nslots - GX_SUB) / (long double) range_d;
    h->kmin = kmin;
    const char *wide = getenv("GX_SLOT_WIDE");
    if (range_d < 4294967295.0 && !(wide && wide[0] == '1')) {
        int sh = 0;
        while (sh < 62 && ratio * (long double) (1ULL << (sh + 1)) < 4294967295.0L) sh++;
        h->mode = 2; h->shift = (unsigned int) sh;
        h->scale = (unsigned long long) (ratio * (long double) (1ULL << sh));
        return;
    }
    h->mode = 1; h->shift = 0;
    const long double sc = (long double) 18446744073709551616.0L * ratio;
    h->scale = sc >= 18446744073709551615.0L ? ~0ULL : (unsigned long long) sc;
}

extern "C" int gx_hash_build(gx_ctx *ctx, const gx_table *inner, int key_col, int n_preds, const gx_pred *preds,
                             int n_payload, const int32_t *payload_cols, int unique, gx_hash **out)
{
    if (!ctx || !inner || !out) return GX_ERR_ARG;
    GX_CHECK_ARG(ctx, key_col >= 0 && key_col < inner->ncols, "hash_build: key column %d out of range", key_col);
    int kt = inner->types[key_col];
    GX_CHECK_ARG(ctx, kt == GX_INT4 || kt == GX_INT8 || kt == GX_DATE, "hash_build: key type %d not hashable here (int4/int8/date only)", kt);
    GX_CHECK_ARG(ctx, n_payload >= 0 && n_payload <= GX_MAX_PAYLOAD, "hash_build: n_payload %d", n_payload);
    gx_build_args a; memset(&a, 0, sizeof(a));
    a.key.data = inner->cols[key_col]; a.key.nulls = inner->nulls[key_col]; a.key.type = kt;
    a.npreds = n_preds; a.n_payload = n_payload; a.nrows = inner->nrows;
    int rc = gx_fill_dpreds(ctx, inner, n_preds, preds, a.preds); if (rc) return rc;
    int bytes = 0;
    gx_hash *h = (gx_hash *) calloc(1, sizeof(gx_hash));
    h->ctx = ctx; h->key_type = kt; h->n_payload = n_payload; h->unique = unique;
    struct hash_guard { gx_hash *h; ~hash_guard() { if (h) gx_hash_free(h); } } guard{h};     // every early return frees the table
    for (int i = 0; i < n_payload; i++) {
        int c = payload_cols[i];
        if (c < 0 || c >= inner->ncols) { GX_SET_ERR(ctx, "hash_build: payload column %d out of range", c); return GX_ERR_ARG; }
        if (inner->nulls[c]) { GX_SET_ERR(ctx, "hash_build: nullable payload column %d not supported in-slot", c); return GX_ERR_ARG; }
        a.payload[i].data = inner->cols[c]; a.payload[i].nulls = nullptr; a.payload[i].type = inner->types[c];
        h->payload_types[i] = inner->types[c];
        bytes += gx_type_size(inner->types[c]);
    }
    if (bytes > 8) { GX_SET_ERR(ctx, "hash_build: payload columns total %d bytes > 8", bytes); return GX_ERR_ARG; }

    int64_t want = inner->nrows + inner->nrows / 2 + 16;       // load factor <= 0.67
    h->nslots = gx_pow2_ceil(want);
    h->special_cap = 1 << 16;
    cudaError_t e = gx_tmp_alloc(ctx, (void **) &h->slots, (size_t) h->nslots * sizeof(gx_slot));
    if (e == cudaSuccess) e = gx_tmp_alloc(ctx, (void **) &h->special_payload, (size_t) h->special_cap * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        GX_SET_ERR(ctx, "hash_build: cudaMalloc of %lld slots failed: %s", (long long) h->nslots, cudaGetErrorString(e));
        return GX_ERR_NOMEM;
    }
    // choose the slot function from a strided key sample (64 K keys): interpolation when the
    // keys span at most GX_INTERP_MAX_SPREAD x their count
    h->mode = 0; h->kmin = 0; h->scale = 0; h->shift = 0; h->amask = 1u;
    bool have_ends = false, counts_read = false;
    { const char *w = getenv("GX_SLOT_WIN"); h->win = w ? (unsigned int) atoi(w) : 31u; }
    {
        const char *fm = getenv("GX_SLOT_MODE");
        int want = fm ? atoi(fm) : -1;
        if (inner->nrows >= 4096 && h->nslots >= 64 * GX_SUB && want != 0) {
            const long long nsample = 65536;
            long long stride = inner->nrows / nsample; if (stride < 1) stride = 1;
            long long nthreads = (inner->nrows + stride - 1) / stride;
            long long init[2] = { 0x7fffffffffffffffLL, -0x7fffffffffffffffLL - 1 };
            GX_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch + 4, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
            { gx_launch_scope ls(ctx, "build_sample"); gx_k_key_range<<<(unsigned) ((nthreads + 255) / 256), 256, 0, ctx->stream>>>(a.key, inner->nrows, stride, ctx->d_scratch + 4); }
            GX_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 4, ctx->d_scratch + 4, 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
            // the column's first and last key ride along: the key-ordered build path below needs them
            {
                const int ksz0 = gx_type_size(kt);
                ctx->h_scratch[20] = ctx->h_scratch[21] = 0;
                GX_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 20, (const char *) inner->cols[key_col], ksz0, cudaMemcpyDeviceToHost, ctx->stream));
                GX_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 21, (const char *) inner->cols[key_col] + (size_t) (inner->nrows - 1) * ksz0, ksz0, cudaMemcpyDeviceToHost, ctx->stream));
                have_ends = true;
            }
            GX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            long long lo = ctx->h_scratch[4], hi = ctx->h_scratch[5];
            if (hi > lo) {
                double range = (double) hi - (double) lo + 1.0;
                if (range <= GX_INTERP_MAX_SPREAD * (double) inner->nrows || want == 1) {
                    // widen the sampled range a little: keys outside it still map (they just wrap)
                    double pad = range / 1024.0 + 64.0;
                    double lo_d = (double) lo - pad, range_d = range + 2 * pad;
                    if (lo_d > -9.0e18 && lo_d + range_d < 9.0e18) {
                        gx_set_interpolation(h, (long long) lo_d, range_d);
                    }
                }
            }
        }
    }
    a.sf.mode = h->mode; a.sf.win = h->win; a.sf.shift = h->shift; a.sf.amask = h->amask; a.sf.kmin = h->kmin; a.sf.scale = h->scale; a.sf.mask = (unsigned long long) h->nslots - 1;
    a.slots = h->slots; a.mask = (unsigned long long) h->nslots - 1;
    a.special = h->special_payload; a.special_cap = h->special_cap;
    a.counters = ctx->d_scratch;
    GX_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 2 * sizeof(long long), ctx->stream));
    // big tables: bucket by sub-table, build each sub-table in shared memory
    const char *force = getenv("GX_BUILD_DIRECT");
    bool bucketed = h->nslots >= 64 * GX_SUB && !(force && force[0] == '1');
    long long nscattered = -1;
    for (int pass = 0; pass < 2 && inner->nrows > 0 && bucketed && nscattered < 0; pass++) {
        // second pass only when the interpolation slot function produced long chains / overflow
        GX_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 2 * sizeof(long long), ctx->stream));
        a.sf.mode = h->mode; a.sf.win = h->win; a.sf.shift = h->shift; a.sf.amask = h->amask; a.sf.kmin = h->kmin; a.sf.scale = h->scale;
        gx_bbuild_args ba; memset(&ba, 0, sizeof(ba));
        ba.nsub = h->nslots / GX_SUB;
        // ---- key-ordered build side + order-preserving slots: no bucketing needed
        const char *nosort = getenv("GX_BUILD_NOSORTED");
        if (h->mode != 0 && !(nosort && nosort[0] == '1')) {
            int *d_flag = (int *) (ctx->d_scratch + 6);
            long long ends[2];
            int ksz = gx_type_size(kt);
            cudaMemsetAsync(d_flag, 0, 2 * sizeof(long long), ctx->stream);          // [6] unsorted flag, [7] overflow
            if (!have_ends) {
                ctx->h_scratch[20] = ctx->h_scratch[21] = 0;
                cudaMemcpyAsync(ctx->h_scratch + 20, (const char *) inner->cols[key_col], ksz, cudaMemcpyDeviceToHost, ctx->stream);
                cudaMemcpyAsync(ctx->h_scratch + 21, (const char *) inner->cols[key_col] + (size_t) (inner->nrows - 1) * ksz, ksz, cudaMemcpyDeviceToHost, ctx->stream);
                GX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                have_ends = true;
            }
            ends[0] = ctx->h_scratch[20]; ends[1] = ctx->h_scratch[21];
            long long kmin = ksz == 4 ? (long long) (int) ends[0] : ends[0], kmax = ksz == 4 ? (long long) (int) ends[1] : ends[1];
            if (kmin == GX_EMPTY_KEY) kmin = kmin + 1;                                  // the reserved key lives in the side list
            double range_d = (double) kmax - (double) kmin + 1.0;
            long long *d_start = nullptr;
            if (range_d >= 1.0 && range_d < 9.0e18 && range_d <= GX_INTERP_MAX_SPREAD * (double) inner->nrows &&
                gx_tmp_alloc(ctx, (void **) &d_start, (size_t) (ba.nsub + 1) * sizeof(long long)) == cudaSuccess) {
                // exact bounds (if the column really is ordered): every key maps below nslots - GX_SUB, monotonically
                const long long save_kmin = h->kmin; const unsigned long long save_scale = h->scale;
                const int save_mode = h->mode; const unsigned int save_shift = h->shift;
                gx_set_interpolation(h, kmin, range_d);
                a.sf.mode = h->mode; a.sf.kmin = h->kmin; a.sf.scale = h->scale; a.sf.shift = h->shift; a.sf.amask = h->amask;
                ba.b = a; ba.overflow = (int *) (ctx->d_scratch + 7); ba.start = d_start; ba.unsorted = d_flag;
                cudaMemsetAsync(ctx->d_scratch + 3, 0, sizeof(long long), ctx->stream);
                cudaMemsetAsync(ctx->d_scratch + 5, 0, sizeof(long long), ctx->stream);
                int pk = 0;
                if (kt == GX_INT8 && a.key.nulls == nullptr && a.npreds == 0) {
                    if (a.n_payload == 0) pk = 2;
                    else if (a.n_payload == 1 && (a.payload[0].type == GX_INT4 || a.payload[0].type == GX_DATE)) pk = 1;
                }
                // int8 key without NULLs whose first key is an ordinary key: bounds by search, order and
                // placement verified by the fill (see sorted_lower_bound)
                const char *nosearch = getenv("GX_NO_BOUNDS_SEARCH");
                const bool searched = kt == GX_INT8 && a.key.nulls == nullptr && ends[0] != GX_EMPTY_KEY && ((uintptr_t) a.key.data & 15) == 0 &&
                                      !(nosearch && nosearch[0] == '1');
                // compact 8-byte slots when the exact key span and the payload fit 32 bits each
                const char *nocompact = getenv("GX_NO_COMPACT");
                const bool small_payload = a.n_payload == 0 ? inner->nrows < 0xffffffffLL : bytes <= 4;
                // (a sub-table's keys must be told apart by 31 bits: its key span is range / #sub-tables)
                const double keys_per_slot = range_d / (double) (h->nslots - GX_SUB);
                bool compact = (pk != 0 || searched) && small_payload && h->mode != 0 && keys_per_slot * GX_SUB < 268435456.0 &&
                               !(nocompact && nocompact[0] == '1');
                h->keys_per_slot = keys_per_slot;
                if (compact && gx_tmp_alloc(ctx, (void **) &h->cslots, (size_t) h->nslots * sizeof(gx_cslot)) != cudaSuccess) { h->cslots = nullptr; compact = false; }
                h->amask = compact ? 3u : 1u; h->cspan = compact ? (unsigned long long) (kmax - kmin) + 1ULL : 0ULL;
                a.sf.amask = h->amask; ba.b = a; ba.cslots = h->cslots;
                static bool sattr = false;
                if (!sattr) {
                    cudaFuncSetAttribute(gx_k_sorted_fill<0, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES);
                    cudaFuncSetAttribute(gx_k_sorted_fill<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES);
                    cudaFuncSetAttribute(gx_k_sorted_fill<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES);
                    cudaFuncSetAttribute(gx_k_sorted_fill<1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES_C);
                    cudaFuncSetAttribute(gx_k_sorted_fill<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES_C);
                    sattr = true;
                }
                if (searched) {
                    static bool fattr = false;
                    if (!fattr) {
                        cudaFuncSetAttribute(gx_k_sorted_fill<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES);
                        cudaFuncSetAttribute(gx_k_sorted_fill<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES_C);
                        cudaFuncSetAttribute(gx_k_sorted_fill<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES);
                        cudaFuncSetAttribute(gx_k_sorted_fill<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES);
                        cudaFuncSetAttribute(gx_k_sorted_fill<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES_C);
                        cudaFuncSetAttribute(gx_k_sorted_fill<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES_C);
                        fattr = true;
                    }
                    { gx_launch_scope ls(ctx, "build_bounds"); gx_k_sorted_bounds_search<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const long long *) a.key.data, inner->nrows, a.sf, ba.nsub, h->nslots, d_start); }
                    gx_launch_scope ls(ctx, "build");
                    if (compact && pk == 1) gx_k_sorted_fill<1, true, true><<<ctx->sm_count * 10, FILL_THREADS, FILL_SMEM_BYTES_C, ctx->stream>>>(ba);
                    else if (compact && pk == 2) gx_k_sorted_fill<2, true, true><<<ctx->sm_count * 10, FILL_THREADS, FILL_SMEM_BYTES_C, ctx->stream>>>(ba);
                    else if (compact) gx_k_sorted_fill<0, true, true><<<ctx->sm_count * 10, FILL_THREADS, FILL_SMEM_BYTES_C, ctx->stream>>>(ba);
                    else if (pk == 1) gx_k_sorted_fill<1, false, true><<<ctx->sm_count * 8, FILL_THREADS, FILL_SMEM_BYTES, ctx->stream>>>(ba);
                    else if (pk == 2) gx_k_sorted_fill<2, false, true><<<ctx->sm_count * 8, FILL_THREADS, FILL_SMEM_BYTES, ctx->stream>>>(ba);
                    else gx_k_sorted_fill<0, false, true><<<ctx->sm_count * 8, FILL_THREADS, FILL_SMEM_BYTES, ctx->stream>>>(ba);
                } else {
                    {
                        gx_launch_scope ls(ctx, "build_bounds");
                        if (kt == GX_INT8 && a.key.nulls == nullptr && ((uintptr_t) a.key.data & 15) == 0)
                            gx_k_sorted_bounds_i8<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const long long *) a.key.data, inner->nrows, a.sf, ba.nsub, d_start, d_flag);
                        else
                            gx_k_sorted_bounds<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(a, ba.nsub, d_start, d_flag);
                    }
                    {
                        gx_launch_scope ls(ctx, "build", 2);
                        gx_k_sorted_special<<<1, 256, 0, ctx->stream>>>(a, d_start, d_flag);
                        const unsigned int fgrid = ctx->sm_count * 8;
                        if (compact && pk == 1) gx_k_sorted_fill<1, true, false><<<ctx->sm_count * 10, FILL_THREADS, FILL_SMEM_BYTES_C, ctx->stream>>>(ba);
                        else if (compact) gx_k_sorted_fill<2, true, false><<<ctx->sm_count * 10, FILL_THREADS, FILL_SMEM_BYTES_C, ctx->stream>>>(ba);
                        else if (pk == 1) gx_k_sorted_fill<1, false, false><<<fgrid, FILL_THREADS, FILL_SMEM_BYTES, ctx->stream>>>(ba);
                        else if (pk == 2) gx_k_sorted_fill<2, false, false><<<fgrid, FILL_THREADS, FILL_SMEM_BYTES, ctx->stream>>>(ba);
                        else gx_k_sorted_fill<0, false, false><<<fgrid, FILL_THREADS, FILL_SMEM_BYTES, ctx->stream>>>(ba);
                    }
                }
                cudaError_t e2 = cudaGetLastError();
                if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 8 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
                if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
                gx_tmp_free(ctx, d_start);
                if (e2 != cudaSuccess) { GX_SET_ERR(ctx, "hash_build: %s", cudaGetErrorString(e2)); return GX_ERR_CUDA; }
                if ((int) ctx->h_scratch[6] == 0) {
                    h->avg_chain = ctx->h_scratch[3] > 0 ? (double) ctx->h_scratch[5] / (double) ctx->h_scratch[3] : 0.0;
                    if ((int) ctx->h_scratch[7] == 0 && h->avg_chain <= 4.0) {
                        nscattered = ctx->h_scratch[3]; h->sorted_build = 1; counts_read = true;
                        if (compact) { gx_tmp_free(ctx, h->slots); h->slots = nullptr; }      // the 16-byte form is made on demand (gx_hash_wide)
                        break;
                    }
                    if (compact) { gx_tmp_free(ctx, h->cslots); h->cslots = nullptr; h->amask = 1u; h->cspan = 0; }
                    h->mode = 0; continue;                                               // clustered keys: rebuild with the mixing hash
                }
                if (compact) { gx_tmp_free(ctx, h->cslots); h->cslots = nullptr; h->amask = 1u; h->cspan = 0; }
                // not in key order: back to the sampled slot function and the bucketing passes
                h->kmin = save_kmin; h->scale = save_scale; h->mode = save_mode; h->shift = save_shift;
                a.sf.mode = h->mode; a.sf.kmin = h->kmin; a.sf.scale = h->scale; a.sf.shift = h->shift; a.sf.amask = h->amask;
                GX_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 2 * sizeof(long long), ctx->stream));
            }
        }
        ba.b = a;
        cudaError_t e = gx_tmp_alloc(ctx, (void **) &ba.cursor, (size_t) ba.nsub * sizeof(unsigned int) + sizeof(int));
        if (e == cudaSuccess) e = gx_tmp_alloc(ctx, (void **) &ba.pairs, (size_t) h->nslots * sizeof(gx_slot));
        if (e != cudaSuccess) {
            gx_tmp_free(ctx, ba.cursor);
            GX_SET_ERR(ctx, "hash_build: cudaMalloc of the bucketing buffers failed: %s", cudaGetErrorString(e));
            return GX_ERR_NOMEM;
        }
        ba.overflow = (int *) (ba.cursor + ba.nsub);
        cudaMemsetAsync(ba.cursor, 0, (size_t) ba.nsub * sizeof(unsigned int) + sizeof(int), ctx->stream);
        cudaMemsetAsync(ctx->d_scratch + 3, 0, sizeof(long long), ctx->stream);
        cudaMemsetAsync(ctx->d_scratch + 5, 0, sizeof(long long), ctx->stream);
        long long ntiles = (inner->nrows + 256 * BSCAT - 1) / (256 * BSCAT), maxb = (long long) ctx->sm_count * 8;
        unsigned grid = (unsigned) (ntiles < maxb ? ntiles : maxb);
        int log2nsub = 0; while (((long long) 1 << log2nsub) < ba.nsub) log2nsub++;
        const char *onepass = getenv("GX_BUILD_ONEPASS");
        bool two_level = log2nsub >= 12 && log2nsub <= 18 && !(onepass && onepass[0] == '1');
        if (two_level) {
            gx_bpart_args pa; memset(&pa, 0, sizeof(pa));
            pa.b = a; pa.nsub = ba.nsub; pa.bits_hi = (log2nsub + 1) / 2; pa.bits_lo = log2nsub - pa.bits_hi;
            pa.pairs = ba.pairs; pa.cursor = ba.cursor; pa.overflow = ba.overflow;
            const long long B1 = (long long) 1 << pa.bits_hi;
            pa.cap1 = (inner->nrows / B1) + (inner->nrows / B1) / 4 + 2 * BP_TILE;
            pa.tiles_per_l1 = (pa.cap1 + BP_TILE - 1) / BP_TILE;
            e = gx_tmp_alloc(ctx, (void **) &pa.l1, (size_t) B1 * pa.cap1 * sizeof(gx_slot));
            if (e == cudaSuccess) e = gx_tmp_alloc(ctx, (void **) &pa.cur1, (size_t) B1 * sizeof(unsigned int));
            if (e == cudaSuccess) {
                cudaMemsetAsync(pa.cur1, 0, (size_t) B1 * sizeof(unsigned int), ctx->stream);
                static bool attr = false;
                if (!attr) {
                    cudaFuncSetAttribute(gx_k_bpart1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bp_smem_bytes(512));
                    cudaFuncSetAttribute(gx_k_bpart2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bp_smem_bytes(512));
                    attr = true;
                }
                gx_launch_scope ls(ctx, "build_scatter", 2);
                gx_k_bpart1<<<ctx->sm_count * 2, BP_THREADS, bp_smem_bytes(1 << pa.bits_hi), ctx->stream>>>(pa);
                gx_k_bpart2<<<ctx->sm_count * 2, BP_THREADS, bp_smem_bytes(1 << pa.bits_lo), ctx->stream>>>(pa);
            }
            gx_tmp_free(ctx, pa.l1); gx_tmp_free(ctx, pa.cur1);
            if (e != cudaSuccess) two_level = false;            // no room for the level-1 buffer: one-pass scatter
        }
        if (!two_level) { gx_launch_scope ls(ctx, "build_scatter"); gx_k_bbuild_scatter<<<grid, 256, 0, ctx->stream>>>(ba); }
        {
            static bool attr = false;
            if (!attr) { cudaFuncSetAttribute(gx_k_bbuild_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, FILL_SMEM_BYTES); attr = true; }
            gx_launch_scope ls(ctx, "build", 2);
            gx_k_bbuild_fill<<<ctx->sm_count * 8, FILL_THREADS, FILL_SMEM_BYTES, ctx->stream>>>(ba);
            gx_k_bbuild_total<<<ctx->sm_count, 256, 0, ctx->stream>>>(ba.cursor, ba.nsub, ctx->d_scratch + 3);
        }
        int h_over = 0;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_scratch + 3, ctx->d_scratch + 3, 3 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h_over, ba.overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        gx_tmp_free(ctx, ba.cursor); gx_tmp_free(ctx, ba.pairs);
        if (e != cudaSuccess) { GX_SET_ERR(ctx, "hash_build: %s", cudaGetErrorString(e)); return GX_ERR_CUDA; }
        h->avg_chain = ctx->h_scratch[3] > 0 ? (double) ctx->h_scratch[5] / (double) ctx->h_scratch[3] : 0.0;
        if (h->mode != 0 && (h_over || h->avg_chain > 4.0)) { h->mode = 0; continue; }   // keys were not as uniform as the sample said
        if (h_over) bucketed = false;                 // a sub-table overflowed (heavy key skew): build it the direct way
        else nscattered = ctx->h_scratch[3];
    }
    a.sf.mode = h->mode; a.sf.win = h->win; a.sf.shift = h->shift; a.sf.amask = h->amask; a.sf.kmin = h->kmin; a.sf.scale = h->scale;
    if (inner->nrows > 0 && !bucketed) {
        GX_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 2 * sizeof(long long), ctx->stream));
        { gx_launch_scope ls(ctx, "build_clear"); gx_k_fill_slots<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(h->slots, h->nslots); }
        gx_launch_scope ls(ctx, "build");
        long long nb = (inner->nrows + 255) / 256;
        long long maxb = (long long) ctx->sm_count * 8;
        gx_k_hash_build<<<(unsigned) (nb < maxb ? nb : maxb), 256, 0, ctx->stream>>>(a);
    } else if (inner->nrows == 0) {
        gx_launch_scope ls(ctx, "build_clear"); gx_k_fill_slots<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(h->slots, h->nslots);
    }
    GX_CUDA(ctx, cudaGetLastError());
    if (!counts_read) {
        GX_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        GX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (nscattered >= 0) ctx->h_scratch[0] = nscattered;
    h->nentries = ctx->h_scratch[0] + ctx->h_scratch[1];
    h->special_count = (int) ctx->h_scratch[1];
    if (h->special_count > h->special_cap) {
        GX_SET_ERR(ctx, "hash_build: %d rows carry key INT64_MIN (side list holds %d)", h->special_count, h->special_cap);
        return GX_ERR_ARG;
    }
    guard.h = nullptr;
    *out = h;
    return GX_OK;
}

// compact table -> 16-byte slots, for the consumers that only read that form.  The full key is
// rebuilt from its 31 stored bits and the slot's position: the interpolation maps slot i back to
// key offset ~ i * keys_per_slot (the entry sits within a chain length and a 32-slot window of its
// home), and the offset congruent to the stored bits nearest to that estimate is the key's.
__global__ void gx_k_expand_slots(const gx_cslot *c, gx_slot *w, long long nslots, long long kmin, double keys_per_slot)
{
    long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += stride) {
        gx_cslot v = c[i];
        longlong2 o;
        o.x = GX_EMPTY_KEY; o.y = 0;
        if (v.d) {
            const long long est = (long long) ((double) i * keys_per_slot);
            long long off = (est & ~0x7FFFFFFFLL) | (long long) (v.d >> 1);
            if (off - est > 0x40000000LL) off -= 0x80000000LL;
            else if (est - off > 0x40000000LL) off += 0x80000000LL;
            o.x = kmin + off; o.y = (long long) (unsigned long long) v.payload;
        }
        ((longlong2 *) w)[i] = o;
    }
}
int gx_hash_wide(gx_ctx *ctx, gx_hash *h)
{
    if (h->slots || !h->cslots) return GX_OK;
    cudaError_t e = gx_tmp_alloc(ctx, (void **) &h->slots, (size_t) h->nslots * sizeof(gx_slot));
    if (e != cudaSuccess) { h->slots = nullptr; GX_SET_ERR(ctx, "hash table: cudaMalloc of %lld slots failed: %s", (long long) h->nslots, cudaGetErrorString(e)); return GX_ERR_NOMEM; }
    gx_launch_scope ls(ctx, "build_expand");
    gx_k_expand_slots<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(h->cslots, h->slots, h->nslots, h->kmin, h->keys_per_slot);
    GX_CUDA(ctx, cudaGetLastError());
    return GX_OK;
}

extern "C" int64_t gx_hash_nentries(const gx_hash *h) { return h ? h->nentries : -1; }
extern "C" int64_t gx_hash_nslots(const gx_hash *h) { return h ? h->nslots : -1; }
extern "C" int gx_hash_info(const gx_hash *h, int *slot_mode, double *avg_chain)
{
    if (!h) return GX_ERR_ARG;
    if (slot_mode) *slot_mode = h->sorted_build ? 2 : (h->mode != 0 ? 1 : 0);
    if (avg_chain) *avg_chain = h->avg_chain;
    return GX_OK;
}
extern "C" void gx_hash_free(gx_hash *h)
{
    if (!h) return;
    gx_tmp_free(h->ctx, h->slots);
    gx_tmp_free(h->ctx, h->cslots);
    gx_tmp_free(h->ctx, h->special_payload);
    free(h);
}

// ------------------------------------------------------------ K3 probe
struct gx_probe_args {
    gx_dcol key;
    int npreds, n_out_outer, n_payload, unique;
    gx_dpred preds[GX_MAX_PREDS];
    long long nrows;
    const gx_slot *slots; unsigned long long mask;
    const gx_cslot *cslots; unsigned long long cspan;       // compact 8-byte form (half the bytes: an L2-resident table more often)
    const unsigned long long *special; int special_count; int _pad2; gx_slotfn sf;
    gx_dcol out_src[GX_MAX_COLS];
    void *out[GX_MAX_COLS];
    uint8_t *out_nulls[GX_MAX_COLS];
    int payload_types[GX_MAX_PAYLOAD];
    long long *cursor;             // global output cursor
    long long out_cap;
    int count_only; int join_type;  // GX_JOIN_*
};

__device__ __forceinline__ void emit_payload(const gx_probe_args &a, long long dst, unsigned long long payload, bool matched);
__device__ __forceinline__ void emit_row(const gx_probe_args &a, long long dst, long long r, unsigned long long payload, bool matched = true)
{
    for (int c = 0; c < a.n_out_outer; c++) {
        switch (a.out_src[c].type) {
            case GX_INT4: case GX_DATE: ((int *) a.out[c])[dst] = ((const int *) a.out_src[c].data)[r]; break;
            case GX_CHAR: ((signed char *) a.out[c])[dst] = ((const signed char *) a.out_src[c].data)[r]; break;
            default: ((long long *) a.out[c])[dst] = ((const long long *) a.out_src[c].data)[r]; break;
        }
        if (a.out_nulls[c]) a.out_nulls[c][dst] = a.out_src[c].nulls ? a.out_src[c].nulls[r] : 0;
    }
    emit_payload(a, dst, payload, matched);
}
__device__ __forceinline__ void emit_payload(const gx_probe_args &a, long long dst, unsigned long long payload, bool matched)
{
    if (a.join_type == GX_JOIN_SEMI || a.join_type == GX_JOIN_ANTI) return;      // the inner side is not part of the target list
    const int npo = a.n_payload ? a.n_payload : 1;
    if (a.join_type == GX_JOIN_LEFT)                                               // hj_NullInnerTupleSlot for an unmatched outer row
        for (int i = 0; i < npo; i++) if (a.out_nulls[a.n_out_outer + i]) a.out_nulls[a.n_out_outer + i][dst] = matched ? 0 : 1;
    if (!matched) payload = 0;
    if (a.n_payload == 0) { ((long long *) a.out[a.n_out_outer])[dst] = (long long) payload; return; }
    int shift = 0;
    for (int i = 0; i < a.n_payload; i++) {
        int t = a.payload_types[i]; void *o = a.out[a.n_out_outer + i];
        switch (t) {
            case GX_INT4: case GX_DATE: ((int *) o)[dst] = (int) (payload >> shift); shift += 32; break;
            case GX_CHAR: ((signed char *) o)[dst] = (signed char) (payload >> shift); shift += 8; break;
            default: ((long long *) o)[dst] = (long long) payload; shift += 64; break;
        }
    }
}

// One pass.  Each thread walks its row's probe sequence; matches are appended
// through a warp-aggregated cursor (join output order is unspecified in the
// reference as well).  count_only sizes the output for non-unique builds.
__global__ void __launch_bounds__(256) gx_k_hash_probe(gx_probe_args a)
{
    long long stride = (long long) gridDim.x * blockDim.x;
    long long nloop = (a.nrows + stride - 1) / stride;
    long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    for (long long it = 0; it < nloop; it++, r += stride) {
        bool ok = r < a.nrows && !gx_is_null(a.key, r);
        if (ok) {
#pragma unroll
            for (int p = 0; p < GX_MAX_PREDS; p++) if (p < a.npreds) ok = ok && gx_eval_pred(a.preds[p], r);
        }
        long long key = ok ? gx_load_int(a.key, r) : 0;
        // walk: the warp iterates until every lane has exhausted its chain
        unsigned long long s = gx_slot_index(key, a.sf);
        bool special = ok && key == GX_EMPTY_KEY;
        int sp_i = 0;
        bool active = ok;
        while (__any_sync(0xffffffffu, active)) {
            bool hit = false; unsigned long long payload = 0;
            if (active) {
                if (special) {
                    if (sp_i < a.special_count) { hit = true; payload = a.special[sp_i++]; if (a.unique) active = false; }
                    else active = false;
                } else {
                    gx_slot sl = a.slots[s];
                    if (sl.key == GX_EMPTY_KEY) active = false;
                    else {
                        if (sl.key == key) { hit = true; payload = sl.payload; if (a.unique) active = false; }
                        s = gx_next_slot(s, a.mask);
                    }
                }
            }
            unsigned int m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                long long base = 0;
                if (lane == 0) base = (long long) atomicAdd((unsigned long long *) a.cursor, (unsigned long long) __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit && !a.count_only) {
                    long long dst = base + __popc(m & ((1u << lane) - 1));
                    if (dst < a.out_cap) emit_row(a, dst, r, payload);
                }
            }
        }
    }
}

// Unique build side (inner_unique, nodeHashjoin.c:859-861): at most one match per outer row.  A warp takes
// tiles of 4 x 32 rows: the four key loads and then the four home-pair loads (one 32-byte sector each) are in
// flight together, matches are numbered with ballots and the warp claims its output range with ONE atomic per
// tile; the surviving rows are copied column by column.
#define PT_K 4
template <bool COMPACT> __global__ void gx_k_hash_probe_unique(gx_probe_args a);
__device__ __forceinline__ void ld_pair(const gx_slot *p, long long &k0, unsigned long long &p0, long long &k1, unsigned long long &p1)
{
    asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(k0), "=l"(p0), "=l"(k1), "=l"(p1) : "l"(p));
}
template <bool COMPACT>
// no min-blocks on purpose: any value (even 1 or 4, which leave the occupancy as it is) changes ptxas' schedule and costs
// 20-35 % here (profiles/r02_occupancy_variants.txt)
__global__ void __launch_bounds__(256) gx_k_hash_probe_unique(gx_probe_args a)
{
    const int lane = threadIdx.x & 31;
    const unsigned int lt = (1u << lane) - 1;
    const long long nwarp = ((long long) gridDim.x * blockDim.x) >> 5, wid = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long tile = 32LL * PT_K;
    for (long long base = wid * tile; base < a.nrows; base += nwarp * tile) {
        long long r[PT_K], key[PT_K]; bool ok[PT_K], hit[PT_K], quals_ok[PT_K]; unsigned long long pay[PT_K], s[PT_K];
#pragma unroll
        for (int j = 0; j < PT_K; j++) { r[j] = base + j * 32 + lane; quals_ok[j] = r[j] < a.nrows; }
        for (int p = 0; p < a.npreds; p++) {
#pragma unroll
            for (int j = 0; j < PT_K; j++) if (quals_ok[j]) quals_ok[j] = gx_eval_pred(a.preds[p], r[j]);
        }
#pragma unroll
        for (int j = 0; j < PT_K; j++) ok[j] = quals_ok[j] && !gx_is_null(a.key, r[j]);      // a NULL key never matches (hashStrict)
#pragma unroll
        for (int j = 0; j < PT_K; j++) key[j] = ok[j] ? gx_load_int(a.key, r[j]) : 0;
        long long k0[PT_K], k1[PT_K]; unsigned long long p0[PT_K], p1[PT_K];
        if (COMPACT) {
            // four 8-byte slots {d, payload} per 32-byte sector; d = ((key - kmin) << 1) | 1 truncated to 32 bits, 0 = empty
            unsigned int d[PT_K];
#pragma unroll
            for (int j = 0; j < PT_K; j++) {
                const bool inspan = ok[j] && ((unsigned long long) key[j] - (unsigned long long) a.sf.kmin) < a.cspan;   // outside the build side's key span: no partner
                s[j] = gx_slot_index(key[j], a.sf);
                d[j] = GX_CSLOT_D(key[j], a.sf.kmin);
                k0[j] = k1[j] = 0; p0[j] = p1[j] = 0;
                if (inspan) ld_pair((const gx_slot *) (a.cslots + s[j]), k0[j], p0[j], k1[j], p1[j]);
                else d[j] = 0xfffffffeu;                               // even: matches no stored d (they are odd), the all-zero group ends the walk
            }
#pragma unroll
            for (int j = 0; j < PT_K; j++) {
                hit[j] = false; pay[j] = 0;
                if (!ok[j]) continue;
                for (;;) {
                    const unsigned int d0 = (unsigned int) k0[j], d1 = (unsigned int) p0[j], d2 = (unsigned int) k1[j], d3 = (unsigned int) p1[j];
                    const unsigned long long m = d0 == d[j] ? (unsigned long long) k0[j] : d1 == d[j] ? p0[j] : d2 == d[j] ? (unsigned long long) k1[j] : p1[j];
                    hit[j] = (d0 == d[j]) | (d1 == d[j]) | (d2 == d[j]) | (d3 == d[j]);
                    pay[j] = m >> 32;
                    if (hit[j] | (d0 == 0u) | (d1 == 0u) | (d2 == 0u) | (d3 == 0u)) break;
                    s[j] = gx_next_quad(s[j], a.mask);
                    ld_pair((const gx_slot *) (a.cslots + s[j]), k0[j], p0[j], k1[j], p1[j]);
                }
            }
        } else {
#pragma unroll
        for (int j = 0; j < PT_K; j++) {
            s[j] = gx_slot_index(key[j], a.sf);            k0[j] = k1[j] = GX_EMPTY_KEY; p0[j] = p1[j] = 0;
            if (ok[j] && key[j] != GX_EMPTY_KEY) ld_pair(a.slots + s[j], k0[j], p0[j], k1[j], p1[j]);
        }
#pragma unroll
        for (int j = 0; j < PT_K; j++) {
            hit[j] = false; pay[j] = 0;
            if (!ok[j]) continue;
            if (key[j] == GX_EMPTY_KEY) { hit[j] = a.special_count > 0; if (hit[j]) pay[j] = a.special[0]; continue; }
            for (;;) {
                if (k0[j] == key[j]) { hit[j] = true; pay[j] = p0[j]; break; }
                if (k0[j] == GX_EMPTY_KEY) break;
                if (k1[j] == key[j]) { hit[j] = true; pay[j] = p1[j]; break; }
                if (k1[j] == GX_EMPTY_KEY) break;
                s[j] = gx_next_pair(s[j], a.mask);
                ld_pair(a.slots + s[j], k0[j], p0[j], k1[j], p1[j]);
            }
        }
        }
        __syncwarp();
        // which rows produce output (nodeHashjoin.c:569-668): INNER/SEMI the matched ones, ANTI the unmatched ones
        // (HJ_FILL_OUTER_TUPLE without a match), LEFT all of them (unmatched with a NULL inner side)
        bool matched[PT_K];
#pragma unroll
        for (int j = 0; j < PT_K; j++) {
            matched[j] = hit[j];
            const bool scanned = quals_ok[j];                   // passed the scan quals (a NULL key included: it just never matches)
            if (a.join_type == GX_JOIN_ANTI) hit[j] = scanned && !matched[j];
            else if (a.join_type == GX_JOIN_LEFT) hit[j] = scanned;
        }
        unsigned int m[PT_K]; int total = 0;
#pragma unroll
        for (int j = 0; j < PT_K; j++) { m[j] = __ballot_sync(0xffffffffu, hit[j]); total += __popc(m[j]); }
        if (total == 0) continue;
        long long dst0 = 0;
        if (lane == 0) dst0 = (long long) atomicAdd((unsigned long long *) a.cursor, (unsigned long long) total);
        dst0 = __shfl_sync(0xffffffffu, dst0, 0);
        if (a.count_only) continue;
        // outer columns: column by column, the tile's loads in flight together; then the inner side
        long long dd[PT_K];
#pragma unroll
        for (int j = 0; j < PT_K; j++) { dd[j] = dst0 + __popc(m[j] & lt); hit[j] = hit[j] && dd[j] < a.out_cap; dst0 += __popc(m[j]); }
        for (int c = 0; c < a.n_out_outer; c++) gx_copy_rows<PT_K>(a.out_src[c], a.out[c], a.out_nulls[c], r, dd, hit);
        if (a.join_type != GX_JOIN_SEMI && a.join_type != GX_JOIN_ANTI) {
#pragma unroll
            for (int j = 0; j < PT_K; j++) if (hit[j]) emit_payload(a, dd[j], pay[j], matched[j]);
        }
    }
}

extern "C" int gx_hash_probe_ex(gx_ctx *ctx, const gx_table *outer, int key_col, int n_preds, const gx_pred *preds,
                                const gx_hash *h, int join_type, int n_out_outer, const int32_t *out_outer_cols, gx_table **out);
extern "C" int gx_hash_probe(gx_ctx *ctx, const gx_table *outer, int key_col, int n_preds, const gx_pred *preds,
                             const gx_hash *h, int n_out_outer, const int32_t *out_outer_cols, gx_table **out)
{
    return gx_hash_probe_ex(ctx, outer, key_col, n_preds, preds, h, GX_JOIN_INNER, n_out_outer, out_outer_cols, out);
}
extern "C" int gx_hash_probe_ex(gx_ctx *ctx, const gx_table *outer, int key_col, int n_preds, const gx_pred *preds,
                                const gx_hash *h, int join_type, int n_out_outer, const int32_t *out_outer_cols, gx_table **out)
{
    if (!ctx || !outer || !h || !out) return GX_ERR_ARG;
    GX_CHECK_ARG(ctx, key_col >= 0 && key_col < outer->ncols, "hash_probe: key column %d out of range", key_col);
    int kt = outer->types[key_col];
    GX_CHECK_ARG(ctx, kt == GX_INT4 || kt == GX_INT8 || kt == GX_DATE, "hash_probe: key type %d not supported", kt);
    GX_CHECK_ARG(ctx, join_type == GX_JOIN_INNER || join_type == GX_JOIN_LEFT || join_type == GX_JOIN_SEMI || join_type == GX_JOIN_ANTI,
                 "hash_probe: join type %d not supported (inner, left, semi, anti)", join_type);
    GX_CHECK_ARG(ctx, join_type != GX_JOIN_LEFT || h->unique, "hash_probe: LEFT JOIN needs a unique build side (inner_unique)");
    const bool no_inner_cols = join_type == GX_JOIN_SEMI || join_type == GX_JOIN_ANTI;
    int n_pay_out = no_inner_cols ? 0 : (h->n_payload ? h->n_payload : 1);
    GX_CHECK_ARG(ctx, n_out_outer >= 0 && n_out_outer + n_pay_out <= GX_MAX_COLS && n_out_outer + n_pay_out > 0, "hash_probe: bad output column count");
    gx_probe_args a; memset(&a, 0, sizeof(a));
    a.key.data = outer->cols[key_col]; a.key.nulls = outer->nulls[key_col]; a.key.type = kt;
    a.npreds = n_preds; a.n_out_outer = n_out_outer; a.n_payload = h->n_payload; a.unique = h->unique; a.join_type = join_type;
    const bool first_match_only_c = h->unique || join_type != GX_JOIN_INNER;
    const bool use_compact = h->cslots != nullptr && first_match_only_c;       // the 8-byte slots are probed as they are
    if (!use_compact) { int wrc = gx_hash_wide(ctx, const_cast<gx_hash *>(h)); if (wrc) return wrc; }
    a.nrows = outer->nrows; a.slots = h->slots; a.mask = (unsigned long long) h->nslots - 1;
    a.cslots = h->cslots; a.cspan = h->cspan;
    a.special = h->special_payload; a.special_count = h->special_count;
    a.sf.mode = h->mode; a.sf.win = h->win; a.sf.shift = h->shift; a.sf.amask = h->amask; a.sf.kmin = h->kmin; a.sf.scale = h->scale; a.sf.mask = (unsigned long long) h->nslots - 1;
    int rc = gx_fill_dpreds(ctx, outer, n_preds, preds, a.preds); if (rc) return rc;
    int32_t types[GX_MAX_COLS]; bool hn[GX_MAX_COLS];
    for (int c = 0; c < n_out_outer; c++) {
        int oc = out_outer_cols[c];
        GX_CHECK_ARG(ctx, oc >= 0 && oc < outer->ncols, "hash_probe: output column %d out of range", oc);
        a.out_src[c].data = outer->cols[oc]; a.out_src[c].nulls = outer->nulls[oc]; a.out_src[c].type = outer->types[oc];
        types[c] = outer->types[oc]; hn[c] = outer->nulls[oc] != nullptr;
    }
    const bool left = join_type == GX_JOIN_LEFT;
    if (!no_inner_cols && h->n_payload == 0) { types[n_out_outer] = GX_INT8; hn[n_out_outer] = left; }
    for (int i = 0; i < h->n_payload; i++) { if (!no_inner_cols) { types[n_out_outer + i] = h->payload_types[i]; hn[n_out_outer + i] = left; } a.payload_types[i] = h->payload_types[i]; }
    a.cursor = ctx->d_scratch + 2;
    long long nb = (outer->nrows + 255) / 256, maxb = (long long) ctx->sm_count * 8;
    unsigned grid = (unsigned) (nb < maxb ? (nb > 0 ? nb : 1) : maxb);
    long long out_cap = outer->nrows;
    const bool first_match_only = h->unique || join_type != GX_JOIN_INNER;      // semi/anti stop at the first match
    if (!first_match_only) {
        // size the output first
        a.count_only = 1; a.out_cap = 0;
        GX_CUDA(ctx, cudaMemsetAsync(a.cursor, 0, sizeof(long long), ctx->stream));
        { gx_launch_scope ls(ctx, "probe_count"); gx_k_hash_probe<<<grid, 256, 0, ctx->stream>>>(a); }
        GX_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, a.cursor, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        GX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        out_cap = ctx->h_scratch[0];
    }
    gx_table *t;
    rc = gx_table_alloc_like(ctx, n_out_outer + n_pay_out, types, hn, out_cap, &t); if (rc) return rc;
    for (int c = 0; c < t->ncols; c++) { a.out[c] = t->cols[c]; a.out_nulls[c] = t->nulls[c]; }
    a.count_only = 0; a.out_cap = out_cap;
    GX_CUDA(ctx, cudaMemsetAsync(a.cursor, 0, sizeof(long long), ctx->stream));
    {
        gx_launch_scope ls(ctx, "probe");
        if (first_match_only) {
            long long nt = (outer->nrows + 32 * PT_K * 8 - 1) / (32 * PT_K * 8);
            const unsigned pgrid = (unsigned) (nt < maxb ? (nt > 0 ? nt : 1) : maxb);
            if (use_compact) gx_k_hash_probe_unique<true><<<pgrid, 256, 0, ctx->stream>>>(a);
            else gx_k_hash_probe_unique<false><<<pgrid, 256, 0, ctx->stream>>>(a);
        } else gx_k_hash_probe<<<grid, 256, 0, ctx->stream>>>(a);
    }
    cudaError_t e = cudaMemcpyAsync(ctx->h_scratch, a.cursor, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { gx_table_free(t); GX_SET_ERR(ctx, "hash_probe: %s", cudaGetErrorString(e)); return GX_ERR_CUDA; }
    t->nrows = ctx->h_scratch[0] < out_cap ? ctx->h_scratch[0] : out_cap;
    *out = t;
    return GX_OK;
}
