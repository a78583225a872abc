/*
 * orc_bloom.c — restatement of the reference's split block bloom filter
 * (src/backend/utils/misc/bloomfilter.c:20-230, scalar path).
 * TEST INFRASTRUCTURE (see otb_oracle.h).
 */
#include <math.h>
#include <stdlib.h>
#include "otb_oracle.h"

#define LOG_BUCKET_BYTE_SIZE 5      /* 32-byte buckets: 8 x uint32 */
#define BUCKET_WORDS 8

struct orc_bloom { int logNumBuckets; uint64_t directoryMask; uint64_t ninsert; uint32_t *directory; };

/* MinLogSpace, bloomfilter.c:20-35: k = 8 hash functions */
static int MinLogSpace(int64_t ndv, double fpp)
{
    const double k = BUCKET_WORDS;
    double m;
    if (ndv <= 0) return 0;
    m = -k * (double) ndv / log(1 - pow(fpp, 1.0 / k));
    int v = (int) ceil(log2(m / 8));
    return v > 0 ? v : 0;
}
/* Rehash32to32, bloomfilter.c:38-52 */
static inline uint32_t Rehash32to32(uint32_t hash)
{
    const uint64_t m = 0x7850f11ec6d14889ull, a = 0x6773610597ca4c63ull;
    return (uint32_t) (((uint64_t) hash * m + a) >> 32);
}
static const uint32_t REHASH[8] = { 0x47b6137bU, 0x44974d91U, 0x8824ad5bU, 0xa2b7289dU,
                                    0x705495c7U, 0x2df1424bU, 0x9efc4947U, 0x5c6bfb31U };

/* BlockBloomFilterInit(nrows, fpp = BLOOM_ERROR_RATE 0.05 at the hash-join call site, nodeHash.c:717) */
orc_bloom *orc_bloom_create(int64_t nrows)
{
    int logMemorySize = MinLogSpace(nrows, 0.05);     /* BLOOM_ERROR_RATE, nodeHash.c:57 */
    int logNumBuckets = logMemorySize - LOG_BUCKET_BYTE_SIZE;
    if (logNumBuckets < 1) logNumBuckets = 1;
    if (logNumBuckets > 20) return NULL;            /* "give up using bloom filter" */
    orc_bloom *b = (orc_bloom *) calloc(1, sizeof(*b));
    b->logNumBuckets = logNumBuckets;
    b->directoryMask = (1ull << logNumBuckets) - 1;
    b->directory = (uint32_t *) calloc((size_t) 1 << (logNumBuckets + 3), sizeof(uint32_t));
    return b;
}
void orc_bloom_insert(orc_bloom *b, uint32_t hash)
{
    uint32_t bucketIdx = (uint32_t) (Rehash32to32(hash) & b->directoryMask);
    for (int i = 0; i < BUCKET_WORDS; i++) {
        uint32_t hval = (REHASH[i] * hash) >> (32 - 5);
        b->directory[(size_t) bucketIdx * 8 + i] |= 1U << hval;
    }
    b->ninsert++;
}
int orc_bloom_find(const orc_bloom *b, uint32_t hash)
{
    if (!b) return 1;
    uint32_t bucketIdx = (uint32_t) (Rehash32to32(hash) & b->directoryMask);
    for (int i = 0; i < BUCKET_WORDS; i++) {
        uint32_t hval = (REHASH[i] * hash) >> (32 - 5);
        if (!(b->directory[(size_t) bucketIdx * 8 + i] & (1U << hval))) return 0;
    }
    return 1;
}
int orc_bloom_log_num_buckets(const orc_bloom *b) { return b ? b->logNumBuckets : -1; }
const uint32_t *orc_bloom_words(const orc_bloom *b, int64_t *nwords)
{
    if (nwords) *nwords = (int64_t) 1 << (b->logNumBuckets + 3);
    return b->directory;
}
void orc_bloom_free(orc_bloom *b) { if (b) { free(b->directory); free(b); } }
