/*
 * orc_gen.c — host-side synthetic TPC-H-shaped tables.
 * TEST INFRASTRUCTURE (see otb_oracle.h).  The recipe itself lives in
 * include/gx_tpch_gen.h; placement on datanodes follows the reference's SHARD
 * rule on the distribution key (l_orderkey / o_orderkey / c_custkey).
 */
#include "otb_oracle.h"
#include "../include/gx_tpch_gen.h"

static int keep_row(const int32_t *shardmap, int type, int64_t key, int node, int nnodes)
{
    if (nnodes <= 1)
        return 1;
    return orc_route_node(shardmap, type, key, 0) == node;
}

int64_t orc_gen_orders(int sf, int64_t order0, int64_t order1, int node, int nnodes,
                       int64_t *o_orderkey, int32_t *o_custkey, int32_t *o_orderdate,
                       int32_t *o_shippriority)
{
    int32_t map[GX_SHARD_MAP_SHARD_NUM];
    int64_t n = 0;
    orc_default_shardmap(map, nnodes > 0 ? nnodes : 1);
    for (int64_t i = order0; i < order1; i++) {
        int64_t key = gxg_o_orderkey(i);
        if (!keep_row(map, GX_INT8, key, node, nnodes))
            continue;
        if (o_orderkey)     o_orderkey[n] = key;
        if (o_custkey)      o_custkey[n] = gxg_o_custkey(i, sf);
        if (o_orderdate)    o_orderdate[n] = gxg_o_orderdate(i);
        if (o_shippriority) o_shippriority[n] = gxg_o_shippriority(i);
        n++;
    }
    return n;
}

int64_t orc_gen_lineitem_count(int sf, int64_t order0, int64_t order1, int node, int nnodes)
{
    int32_t map[GX_SHARD_MAP_SHARD_NUM];
    int64_t n = 0;
    (void) sf;
    orc_default_shardmap(map, nnodes > 0 ? nnodes : 1);
    for (int64_t i = order0; i < order1; i++)
        if (keep_row(map, GX_INT8, gxg_o_orderkey(i), node, nnodes))
            n += gxg_l_nlines(i);
    return n;
}

int64_t orc_gen_lineitem(int sf, int64_t order0, int64_t order1, int node, int nnodes,
                         int64_t *l_orderkey, double *l_quantity, double *l_extendedprice,
                         double *l_discount, double *l_tax, int32_t *l_shipdate,
                         int8_t *l_returnflag, int8_t *l_linestatus)
{
    int32_t map[GX_SHARD_MAP_SHARD_NUM];
    int64_t n = 0;
    orc_default_shardmap(map, nnodes > 0 ? nnodes : 1);
    for (int64_t i = order0; i < order1; i++) {
        int64_t key = gxg_o_orderkey(i);
        if (!keep_row(map, GX_INT8, key, node, nnodes))
            continue;
        int32_t nl = gxg_l_nlines(i);
        for (int32_t j = 0; j < nl; j++) {
            if (l_orderkey)      l_orderkey[n] = key;
            if (l_quantity)      l_quantity[n] = gxg_l_quantity(i, j);
            if (l_extendedprice) l_extendedprice[n] = gxg_l_extendedprice(i, j, sf);
            if (l_discount)      l_discount[n] = gxg_l_discount(i, j);
            if (l_tax)           l_tax[n] = gxg_l_tax(i, j);
            if (l_shipdate)      l_shipdate[n] = gxg_l_shipdate(i, j);
            if (l_returnflag)    l_returnflag[n] = gxg_l_returnflag(i, j);
            if (l_linestatus)    l_linestatus[n] = gxg_l_linestatus(i, j);
            n++;
        }
    }
    return n;
}

int64_t orc_gen_customer(int sf, int64_t c0, int64_t c1, int node, int nnodes,
                         int32_t *c_custkey, int8_t *c_mktsegment)
{
    int32_t map[GX_SHARD_MAP_SHARD_NUM];
    int64_t n = 0;
    (void) sf;
    orc_default_shardmap(map, nnodes > 0 ? nnodes : 1);
    for (int64_t i = c0; i < c1; i++) {
        int32_t key = gxg_c_custkey(i);
        if (!keep_row(map, GX_INT4, key, node, nnodes))
            continue;
        if (c_custkey)    c_custkey[n] = key;
        if (c_mktsegment) c_mktsegment[n] = gxg_c_mktsegment(i);
        n++;
    }
    return n;
}
