/*
 * gx_tpch_gen.h — the synthetic TPC-H-shaped data recipe.
 *
 * Pure integer arithmetic plus one correctly-rounded IEEE division per float8
 * value, so the host (oracle/, plain C) and the device (csrc/gen.cu) produce
 * bit-identical tables from the same (table, row, column) coordinates.  This
 * is not reference behaviour — the reference ships no data generator — it is
 * the shared definition of the benchmark input.  Included by both sides; a
 * test (tests/test_generator.py) checks device output == host output.
 *
 * Dates are DateADT: int32 days since 2000-01-01 (utils/date.h in the reference).
 */
#ifndef GX_TPCH_GEN_H
#define GX_TPCH_GEN_H

#include <stdint.h>

#if defined(__CUDACC__)
#define GXG_HD __host__ __device__ __forceinline__
#else
#define GXG_HD static inline
#endif

#define GXG_SEED            20240922ULL
#define GXG_T_ORDERS        1
#define GXG_T_LINEITEM      2
#define GXG_T_CUSTOMER      3

#define GXG_DATE_1992_01_01 (-2922)
#define GXG_DATE_1995_06_17 (-1659)
#define GXG_DATE_1998_08_02 (-517)
#define GXG_NDATES          2406      /* 1992-01-01 .. 1998-08-02 inclusive */

/* fixed schemas (column numbers used by plans, tests and bench.py) */
/* orders   */ enum { GXG_O_ORDERKEY = 0, GXG_O_CUSTKEY = 1, GXG_O_ORDERDATE = 2,
                      GXG_O_SHIPPRIORITY = 3, GXG_O_NCOLS = 4 };
/* lineitem */ enum { GXG_L_ORDERKEY = 0, GXG_L_QUANTITY = 1, GXG_L_EXTENDEDPRICE = 2,
                      GXG_L_DISCOUNT = 3, GXG_L_TAX = 4, GXG_L_SHIPDATE = 5,
                      GXG_L_RETURNFLAG = 6, GXG_L_LINESTATUS = 7, GXG_L_NCOLS = 8 };
/* customer */ enum { GXG_C_CUSTKEY = 0, GXG_C_MKTSEGMENT = 1, GXG_C_NCOLS = 2 };

GXG_HD uint64_t gxg_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

GXG_HD uint64_t gxg_rng(uint32_t table, uint64_t row, uint32_t col)
{
    return gxg_splitmix64(GXG_SEED ^ ((uint64_t) table << 56) ^ (row << 8) ^ (uint64_t) col);
}

GXG_HD int64_t gxg_norders(int sf)    { return 1500000LL * sf; }
GXG_HD int64_t gxg_ncustomers(int sf) { return 150000LL * sf; }

/* ---- orders ---- */
GXG_HD int64_t gxg_o_orderkey(int64_t i)
{
    /* TPC-H sparse keys: 8 of every 32 used */
    return (int64_t) ((((uint64_t) i >> 3) << 5) | ((uint64_t) i & 7)) + 1;
}
GXG_HD int32_t gxg_o_custkey(int64_t i, int sf)
{
    /* uniform over [1, 150k*sf] excluding multiples of 3 */
    uint64_t C = (uint64_t) gxg_ncustomers(sf);
    uint64_t nvalid = C - C / 3;
    uint64_t u = gxg_rng(GXG_T_ORDERS, (uint64_t) i, 1) % nvalid;
    return (int32_t) (u + u / 2 + 1);
}
GXG_HD int32_t gxg_o_orderdate(int64_t i)
{
    return GXG_DATE_1992_01_01 + (int32_t) (gxg_rng(GXG_T_ORDERS, (uint64_t) i, 2) % GXG_NDATES);
}
GXG_HD int32_t gxg_o_shippriority(int64_t i) { (void) i; return 0; }

/* ---- lineitem: lines (i, j), j in [0, nlines(i)) ---- */
GXG_HD int32_t gxg_l_nlines(int64_t i)
{
    return 1 + (int32_t) (gxg_rng(GXG_T_LINEITEM, (uint64_t) i << 3, 15) % 7);
}
GXG_HD uint64_t gxg_l_row(int64_t i, int32_t j) { return ((uint64_t) i << 3) | (uint64_t) j; }

GXG_HD int32_t gxg_l_quantity_i(int64_t i, int32_t j)
{
    return 1 + (int32_t) (gxg_rng(GXG_T_LINEITEM, gxg_l_row(i, j), 1) % 50);
}
GXG_HD int64_t gxg_l_extendedprice_cents(int64_t i, int32_t j, int sf)
{
    uint64_t partkey = 1 + gxg_rng(GXG_T_LINEITEM, gxg_l_row(i, j), 2) % (200000ULL * (uint64_t) sf);
    int64_t price_cents = 90000 + (int64_t) ((partkey / 10) % 20001) + 100 * (int64_t) (partkey % 1000);
    return (int64_t) gxg_l_quantity_i(i, j) * price_cents;
}
GXG_HD double gxg_l_quantity(int64_t i, int32_t j) { return (double) gxg_l_quantity_i(i, j); }
GXG_HD double gxg_l_extendedprice(int64_t i, int32_t j, int sf)
{
    return (double) gxg_l_extendedprice_cents(i, j, sf) / 100.0;
}
GXG_HD double gxg_l_discount(int64_t i, int32_t j)
{
    return (double) (gxg_rng(GXG_T_LINEITEM, gxg_l_row(i, j), 3) % 11) / 100.0;
}
GXG_HD double gxg_l_tax(int64_t i, int32_t j)
{
    return (double) (gxg_rng(GXG_T_LINEITEM, gxg_l_row(i, j), 4) % 9) / 100.0;
}
GXG_HD int32_t gxg_l_shipdate(int64_t i, int32_t j)
{
    return gxg_o_orderdate(i) + 1 + (int32_t) (gxg_rng(GXG_T_LINEITEM, gxg_l_row(i, j), 5) % 121);
}
GXG_HD int8_t gxg_l_returnflag(int64_t i, int32_t j)
{
    int32_t receipt = gxg_l_shipdate(i, j) + 1 +
                      (int32_t) (gxg_rng(GXG_T_LINEITEM, gxg_l_row(i, j), 6) % 30);
    if (receipt <= GXG_DATE_1995_06_17)
        return (gxg_rng(GXG_T_LINEITEM, gxg_l_row(i, j), 7) & 1) ? 'R' : 'A';
    return 'N';
}
GXG_HD int8_t gxg_l_linestatus(int64_t i, int32_t j)
{
    return gxg_l_shipdate(i, j) > GXG_DATE_1995_06_17 ? 'O' : 'F';
}

/* ---- customer ---- */
GXG_HD int32_t gxg_c_custkey(int64_t i) { return (int32_t) (i + 1); }
GXG_HD int8_t gxg_c_mktsegment(int64_t i)
{
    /* AUTOMOBILE, BUILDING, FURNITURE, HOUSEHOLD, MACHINERY by initial */
    const char seg[5] = { 'A', 'B', 'F', 'H', 'M' };
    return seg[gxg_rng(GXG_T_CUSTOMER, (uint64_t) i, 1) % 5];
}

#endif /* GX_TPCH_GEN_H */
