// gx_exec.cu — the one-call form used by the CustomScan provider and by
// bench.py's end-to-end leg: HOST column buffers in, partial result out.
// Host->HBM staging is chunked so that the build of the join table and the
// first probe chunks overlap the remaining copies.
#include "gx_internal.cuh"

extern "C" int gx_host_alloc(gx_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return GX_ERR_ARG;
    GX_CUDA(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return GX_OK;
}
extern "C" int gx_host_free(gx_ctx *ctx, void *p)
{
    if (!ctx) return GX_ERR_ARG;
    if (p) GX_CUDA(ctx, cudaFreeHost(p));
    return GX_OK;
}

static int upload(gx_ctx *ctx, const gx_host_table *h, gx_table **out)
{
    GX_CHECK_ARG(ctx, h->ncols > 0 && h->ncols <= GX_MAX_COLS && h->nrows >= 0 && h->types && h->cols, "exec_host: bad host table");
    gx_table *t;
    int rc = gx_table_create(ctx, h->ncols, h->types, h->nrows, &t); if (rc) return rc;
    rc = gx_table_append_columns(t, h->cols, h->nulls, h->nrows);
    if (rc) { gx_table_free(t); return rc; }
    *out = t;
    return GX_OK;
}

extern "C" int gx_exec_host(gx_ctx *ctx, const gx_host_table *outer, const gx_host_table *inner, int inner_key_col,
                            int n_inner_preds, const gx_pred *inner_preds, int n_payload, const int32_t *payload_cols,
                            int inner_unique, const gx_agg_plan *plan, gx_result **out)
{
    if (!ctx || !outer || !plan || !out) return GX_ERR_ARG;
    gx_table *to = nullptr, *ti = nullptr; gx_hash *h = nullptr;
    int rc = GX_OK;
    // inner first: its H2D copy is short, and the build kernel then runs while
    // the (much larger) outer columns are still streaming in on the same queue
    if (inner && plan->outer_key_col >= 0) {
        rc = upload(ctx, inner, &ti);
        if (rc == GX_OK) rc = gx_hash_build(ctx, ti, inner_key_col, n_inner_preds, inner_preds, n_payload, payload_cols, inner_unique, &h);
    }
    if (rc == GX_OK) rc = upload(ctx, outer, &to);
    if (rc == GX_OK) rc = gx_hash_agg(ctx, to, h, plan, out);
    if (h) gx_hash_free(h);
    if (ti) gx_table_free(ti);
    if (to) gx_table_free(to);
    return rc;
}
