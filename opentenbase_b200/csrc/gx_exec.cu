// gx_exec.cu — the one-call form used by the CustomScan provider and by
// bench.py's end-to-end leg: HOST column buffers in, partial result out.
// Host->HBM staging is chunked so that the build of the join table and the
// first probe chunks overlap the remaining copies.
#include "gx_internal.cuh"
#include <unistd.h>
#include <sys/syscall.h>

// Pinned staging memory is placed on the NUMA node the GPU hangs off.  On a two-socket host a
// pinned buffer that happens to be allocated on the other socket is copied at a third of the
// PCIe rate (measured through gx_exec_host on B200 hosts: 15-18 GB/s instead of 48 GB/s), and
// which socket a backend's first touch lands on is a coin toss.  The node comes from sysfs
// (/sys/bus/pci/devices/<bus id>/numa_node); the allocation runs under a temporary
// MPOL_BIND policy (raw syscall, no libnuma), retried unbound if the node cannot hold it.  Anything that fails leaves the default policy.
static int gpu_numa_node(int device)
{
    char bus[32] = { 0 }, path[128];
    if (cudaDeviceGetPCIBusId(bus, (int) sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char) (*c - 'A' + 'a');
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
extern "C" int gx_host_alloc(gx_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return GX_ERR_ARG;
    const char *off = getenv("GX_NO_NUMA_BIND");
    int node = (off && off[0] == '1') ? -1 : gpu_numa_node(ctx->device);
    bool bound = false;
    if (node >= 0 && node < 1024) {
        unsigned long mask[16] = { 0 };
        mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
        // MPOL_BIND: "preferred" silently spills to the other node when the local one is full of page cache
        bound = syscall(SYS_set_mempolicy, 2 /* MPOL_BIND */, mask, (unsigned long) (8 * sizeof(mask))) == 0;
    }
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (bound) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, (unsigned long *) nullptr, 0UL);
    if (e != cudaSuccess && bound) { cudaGetLastError(); e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault); }   // node full: anywhere
    GX_CUDA(ctx, e);
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
