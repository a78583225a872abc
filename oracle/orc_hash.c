/*
 * orc_hash.c — restatement of the reference's hash functions and SHARD routing.
 * TEST INFRASTRUCTURE (see otb_oracle.h).  Little-endian hosts only, like the
 * !WORDS_BIGENDIAN paths it follows.
 */
#include <string.h>
#include <stdlib.h>
#include "otb_oracle.h"

#define ROT(x, k) (((x) << (k)) | ((x) >> (32 - (k))))

/* mix(), src/backend/access/hash/hashfunc.c:560-568 */
#define MIX(a, b, c) do { \
    a -= c; a ^= ROT(c, 4);  c += b; \
    b -= a; b ^= ROT(a, 6);  a += c; \
    c -= b; c ^= ROT(b, 8);  b += a; \
    a -= c; a ^= ROT(c, 16); c += b; \
    b -= a; b ^= ROT(a, 19); a += c; \
    c -= b; c ^= ROT(b, 4);  b += a; \
} while (0)

/* final(), hashfunc.c:593-602 */
#define FINAL(a, b, c) do { \
    c ^= b; c -= ROT(b, 14); \
    a ^= c; a -= ROT(c, 11); \
    b ^= a; b -= ROT(a, 25); \
    c ^= b; c -= ROT(b, 16); \
    a ^= c; a -= ROT(c, 4);  \
    b ^= a; b -= ROT(a, 14); \
    c ^= b; c -= ROT(b, 24); \
} while (0)

/* hash_any(), hashfunc.c:619-815: the non-aligned little-endian path is
 * byte-for-byte equivalent to the aligned one, so only it is restated. */
uint32_t orc_hash_any(const unsigned char *k, int keylen)
{
    uint32_t a, b, c, len = (uint32_t) keylen;

    a = b = c = 0x9e3779b9u + len + 3923095u;
    while (len >= 12) {
        a += (k[0] + ((uint32_t) k[1] << 8) + ((uint32_t) k[2] << 16) + ((uint32_t) k[3] << 24));
        b += (k[4] + ((uint32_t) k[5] << 8) + ((uint32_t) k[6] << 16) + ((uint32_t) k[7] << 24));
        c += (k[8] + ((uint32_t) k[9] << 8) + ((uint32_t) k[10] << 16) + ((uint32_t) k[11] << 24));
        MIX(a, b, c);
        k += 12;
        len -= 12;
    }
    switch (len) {              /* all cases fall through; low byte of c is the length's */
        case 11: c += ((uint32_t) k[10] << 24);  /* FALLTHROUGH */
        case 10: c += ((uint32_t) k[9] << 16);   /* FALLTHROUGH */
        case 9:  c += ((uint32_t) k[8] << 8);    /* FALLTHROUGH */
        case 8:  b += ((uint32_t) k[7] << 24);   /* FALLTHROUGH */
        case 7:  b += ((uint32_t) k[6] << 16);   /* FALLTHROUGH */
        case 6:  b += ((uint32_t) k[5] << 8);    /* FALLTHROUGH */
        case 5:  b += k[4];                      /* FALLTHROUGH */
        case 4:  a += ((uint32_t) k[3] << 24);   /* FALLTHROUGH */
        case 3:  a += ((uint32_t) k[2] << 16);   /* FALLTHROUGH */
        case 2:  a += ((uint32_t) k[1] << 8);    /* FALLTHROUGH */
        case 1:  a += k[0];
    }
    FINAL(a, b, c);
    return c;
}

/* hash_uint32(), hashfunc.c:1044-1058 */
uint32_t orc_hash_uint32(uint32_t k)
{
    uint32_t a, b, c;
    a = b = c = 0x9e3779b9u + (uint32_t) sizeof(uint32_t) + 3923095u;
    a += k;
    FINAL(a, b, c);
    return c;
}

/* hashint4(), hashfunc.c:80-84 */
uint32_t orc_hashint4(int32_t v) { return orc_hash_uint32((uint32_t) v); }

/* hashchar(), hashfunc.c:48-52: (int32) of a C char (signed on x86-64) */
uint32_t orc_hashchar(int8_t v) { return orc_hash_uint32((uint32_t) (int32_t) v); }

/* hashint8(), hashfunc.c:92-110: fold hi into lo so equal int4/int8 collide */
uint32_t orc_hashint8(int64_t val)
{
    uint32_t lohalf = (uint32_t) val;
    uint32_t hihalf = (uint32_t) (val >> 32);
    lohalf ^= (val >= 0) ? hihalf : ~hihalf;
    return orc_hash_uint32(lohalf);
}

/* hashfloat8(), hashfunc.c:325-341: -0 and +0 hash alike */
uint32_t orc_hashfloat8(double key)
{
    if (key == (double) 0)
        return 0;
    return orc_hash_any((const unsigned char *) &key, (int) sizeof(key));
}

/* CRC-32C (Castagnoli), reflected polynomial 0x82F63B78; the reference's
 * slicing-by-8 tables (src/port/pg_crc32c_sb8.c) compute the same function.
 * The table is derived from the polynomial here, not copied. */
static uint32_t crc_table[256];
static int crc_table_ready;
static void crc_init(void)
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t r = i;
        for (int j = 0; j < 8; j++)
            r = (r & 1) ? (r >> 1) ^ 0x82F63B78u : (r >> 1);
        crc_table[i] = r;
    }
    crc_table_ready = 1;
}
uint32_t orc_crc32c(uint32_t crc, const void *data, size_t len)
{
    const unsigned char *p = (const unsigned char *) data;
    if (!crc_table_ready)
        crc_init();
    while (len--)
        crc = crc_table[(crc ^ *p++) & 0xFF] ^ (crc >> 8);
    return crc;
}

/* hash_any_new(), hashfunc.c:112-122: INIT_CRC32C/COMP/FIN (port/pg_crc32c.h) */
uint32_t orc_hash_any_new(const unsigned char *k, int keylen)
{
    return orc_crc32c(0xFFFFFFFFu, k, (size_t) keylen) ^ 0xFFFFFFFFu;
}
/* hashint4new(), hashfunc.c:150-156: value widened to int64 first */
uint32_t orc_hashint4new(int32_t v)
{
    int64_t val = (int64_t) v;
    return orc_hash_any_new((const unsigned char *) &val, (int) sizeof(val));
}
/* hashint8new(), hashfunc.c:168-174 */
uint32_t orc_hashint8new(int64_t val)
{
    return orc_hash_any_new((const unsigned char *) &val, (int) sizeof(val));
}
/* hashcharnew() -> hash_uint32_new((int32) char), hashfunc.c:124-140:
 * (uint32) k widened to int64 — note: ZERO-extended, unlike hashint4new */
uint32_t orc_hashcharnew(int8_t v)
{
    uint32_t k = (uint32_t) (int32_t) v;
    int64_t val = (int64_t) k;
    return orc_hash_any_new((const unsigned char *) &val, (int) sizeof(val));
}
/* hashfloat8new(), hashfunc.c:231-244 */
uint32_t orc_hashfloat8new(double key)
{
    if (key == (double) 0)
        return 0;
    return orc_hash_any_new((const unsigned char *) &key, (int) sizeof(key));
}

/* murmurhash32(), src/include/utils/hashutils.h:39-49 */
uint32_t orc_murmurhash32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
/* hash_combine(), hashutils.h:16-21 */
uint32_t orc_hash_combine(uint32_t a, uint32_t b)
{
    a ^= b + 0x9e3779b9u + (a << 6) + (a >> 2);
    return a;
}

static double datum_as_double(int64_t d)
{
    double x;
    memcpy(&x, &d, sizeof(x));
    return x;
}

/* the pg_amproc hash support function of each type (Jenkins family), as
 * looked up by execTuplesHashPrepare for AGG_HASHED and by compute_hash()
 * (hashfunc.c:1094-1240) for SHARD routing: DATE -> hashint4 */
uint32_t orc_hash_datum(int type, int64_t d)
{
    switch (type) {
        case GX_INT4:
        case GX_DATE:   return orc_hashint4((int32_t) d);
        case GX_INT8:   return orc_hashint8(d);
        case GX_CHAR:   return orc_hashchar((int8_t) d);
        case GX_FLOAT8: return orc_hashfloat8(datum_as_double(d));
    }
    abort();
}
/* the HashFuncAssign() substitutes (nodeHash.c:494-567, enable_newhash=true) */
uint32_t orc_hash_datum_new(int type, int64_t d)
{
    switch (type) {
        case GX_INT4:
        case GX_DATE:   return orc_hashint4new((int32_t) d);
        case GX_INT8:   return orc_hashint8new(d);
        case GX_CHAR:   return orc_hashcharnew((int8_t) d);
        case GX_FLOAT8: return orc_hashfloat8new(datum_as_double(d));
    }
    abort();
}

/* EvaluateHashkey(), pgxc/locator/locator.c:1611-1628: rotate-left-1 BEFORE
 * each column (also the first), XOR the column hash unless NULL. */
uint32_t orc_evaluate_hashkey(const int *types, const uint8_t *isnull,
                              const int64_t *datums, int natts)
{
    uint32_t hashkey = 0;
    for (int i = 0; i < natts; i++) {
        hashkey = (hashkey << 1) | ((hashkey & 0x80000000u) ? 1 : 0);
        if (!(isnull && isnull[i]))
            hashkey ^= orc_hash_datum(types[i], datums[i]);
    }
    return hashkey;
}

/* GetNodeIndexByHashValue(), pgxc/shard/shardmap.c:1147-1160:
 * shardIdx = abs((int) hashvalue) % shmemNumShards (4096).  abs(INT_MIN) is
 * INT_MIN with gcc/glibc on x86-64 and INT_MIN % 4096 == 0; written so here. */
int32_t orc_shard_index(uint32_t hashvalue)
{
    int32_t h = (int32_t) hashvalue;
    uint32_t mag = (h < 0) ? (0u - (uint32_t) h) : (uint32_t) h;  /* |h| mod 2^32 */
    int32_t a = (int32_t) mag;            /* INT_MIN stays INT_MIN, like abs() */
    return a % GX_SHARD_MAP_SHARD_NUM;    /* INT_MIN % 4096 == 0 */
}
/* default shard map, catalog/pgxc_shard_map.c:90-96: shard i -> nodes[i % nNodes] */
void orc_default_shardmap(int32_t *map, int nnodes)
{
    for (int i = 0; i < GX_SHARD_MAP_SHARD_NUM; i++)
        map[i] = i % nnodes;
}
int32_t orc_route_node(const int32_t *shardmap, int type, int64_t datum, int isnull)
{
    uint8_t n = (uint8_t) (isnull != 0);
    uint32_t h = orc_evaluate_hashkey(&type, &n, &datum, 1);
    return shardmap[orc_shard_index(h)];
}
