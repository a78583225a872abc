/*
 * gpuexec_provider.c — OpenTenBase CustomScan provider for libgpuexec.so.
 *
 * The host side of the drop-in boundary, in the reference's own
 * language (C) against the reference's own headers.  It plugs the B200 path in
 * where the planner would otherwise put
 *
 *      HashAggregate <- [Hash Join <- Seq Scan, Hash <- Seq Scan]
 *
 * on a datanode: create_upper_paths_hook (optimizer/planner.h:52-56) offers a
 * CustomPath for UPPERREL_GROUP_AGG; the CustomScan is shipped to the datanodes
 * by name (_outCustomScan / _readCustomScan, nodes/outfuncs.c:1288,
 * readfuncs.c:3473) — so this library must be in shared_preload_libraries on
 * the CN and on every DN; the executor methods below stage the relations'
 * visible tuples into HBM page by page (heapgetpage, access/heapam.h:126, one
 * visibility pass per 8 KB page — MVCC/GTS stays on the host) and hand the rest
 * to libgpuexec.so through the C ABI of include/gpuexec.h.  Nothing here throws
 * across the boundary: every gx_* status is turned into ereport(ERROR) on this
 * side, and every handle is released from a ResourceOwner callback because
 * ereport is siglongjmp (utils/elog.h:386-398).
 *
 * Build: against an OpenTenBase install (pg_config --includedir-server), or in
 * this repository against /root/reference/src/include with the stub
 * pg_config.h of oracle/ref (compile check only; `make -C opentenbase_b200/provider`).
 *
 * Status: the executor side (Begin/Exec/End/ReScan/Explain), the plan
 * (de)serialisation into custom_private and the heap-page loader are complete;
 * the planner hook recognises the plan shapes of BASELINE configs 1-3
 * (single-relation or two-relation inner equi-join on one int4/int8 key, Var
 * group keys, count(*) / sum / avg / min / max over float8 Var/Const
 * arithmetic).  It has been compiled against the reference headers but never
 * run inside a live backend (none can be built here: SURVEY.md §8c).
 */
#include "postgres.h"

#include "access/heapam.h"
#include "access/htup_details.h"
#include "access/relscan.h"
#include "catalog/pg_type.h"
#include "commands/explain.h"
#include "executor/executor.h"
#include "fmgr.h"
#include "miscadmin.h"
#include "nodes/extensible.h"
#include "nodes/makefuncs.h"
#include "nodes/nodeFuncs.h"
#include "optimizer/pathnode.h"
#include "optimizer/paths.h"
#include "optimizer/planner.h"
#include "optimizer/tlist.h"
#include "utils/fmgroids.h"
#include "storage/bufmgr.h"
#include "utils/guc.h"
#include "utils/memutils.h"
#include "utils/rel.h"
#include "utils/resowner.h"
#include "utils/snapmgr.h"

#include "gpuexec.h"

PG_MODULE_MAGIC;

void		_PG_init(void);

static bool gpuexec_enabled = true;
static int	gpuexec_device = 0;
static create_upper_paths_hook_type prev_upper_paths_hook = NULL;

/* one CUDA context per backend process, created lazily (never in the postmaster:
 * fork after CUDA initialisation is illegal) */
static gx_ctx *backend_ctx = NULL;

#define GPUEXEC_NAME "GpuExecHashAgg"

/* ---------------------------------------------------------------- state */
typedef struct GpuRelInfo
{
	Index		rti;			/* range-table index of the scanned relation */
	int			ncols;			/* referenced attributes */
	int			attnums[GX_MAX_COLS];	/* 0-based attribute numbers */
	int32		types[GX_MAX_COLS];		/* GX_* */
} GpuRelInfo;

typedef struct GpuExecState
{
	CustomScanState css;
	/* decoded from custom_private */
	GpuRelInfo	outer,
				inner;
	bool		has_join;
	int			inner_key_col;
	int			n_payload;
	int32		payload_cols[GX_MAX_PAYLOAD];
	bool		inner_unique;
	gx_agg_plan plan;
	/* run time */
	Relation	outer_rel,
				inner_rel;
	gx_table   *outer_tab,
			   *inner_tab;
	gx_hash    *hash;
	gx_result  *result;
	int64		ngroups,
				next;
	int64	   *keys;
	double	   *aggs;
	uint8	   *nulls;
	double		load_ms,
				exec_ms;
	bool		done_exec;
} GpuExecState;

/* handles that must die with the query even when ereport() longjmps past us */
typedef struct GpuHandles
{
	struct GpuHandles *next;
	ResourceOwner owner;
	GpuExecState *state;
} GpuHandles;
static GpuHandles *live_handles = NULL;

static void
gpuexec_release_state(GpuExecState *st)
{
	if (st->result)
		gx_result_free(st->result);
	if (st->hash)
		gx_hash_free(st->hash);
	if (st->outer_tab)
		gx_table_free(st->outer_tab);
	if (st->inner_tab)
		gx_table_free(st->inner_tab);
	st->result = NULL;
	st->hash = NULL;
	st->outer_tab = st->inner_tab = NULL;
}

static void
gpuexec_resowner_callback(ResourceReleasePhase phase, bool isCommit, bool isTopLevel, void *arg)
{
	GpuHandles **p = &live_handles;

	if (phase != RESOURCE_RELEASE_AFTER_LOCKS)
		return;
	while (*p)
	{
		GpuHandles *h = *p;

		if (h->owner == CurrentResourceOwner)
		{
			gpuexec_release_state(h->state);
			*p = h->next;
			free(h);
		}
		else
			p = &h->next;
	}
}

#define GX_CHECK(call) \
	do { \
		int			gx_status__ = (call); \
		if (gx_status__ != GX_OK) \
			ereport(ERROR, \
					(errcode(gx_status__ == GX_ERR_OVERFLOW ? ERRCODE_NUMERIC_VALUE_OUT_OF_RANGE : ERRCODE_EXTERNAL_ROUTINE_EXCEPTION), \
					 errmsg("gpuexec: %s", gx_last_error(backend_ctx)))); \
	} while (0)

static void
gpuexec_ensure_context(void)
{
	if (backend_ctx == NULL)
	{
		int			st = gx_init(gpuexec_device, &backend_ctx);

		if (st != GX_OK)
			ereport(ERROR,
					(errcode(ERRCODE_EXTERNAL_ROUTINE_EXCEPTION),
					 errmsg("gpuexec: %s", gx_last_error(NULL))));
	}
}

/* --------------------------------------------------- type / expr mapping */
static int32
gpuexec_type_of(Oid typid)
{
	switch (typid)
	{
		case INT4OID:
			return GX_INT4;
		case INT8OID:
			return GX_INT8;
		case FLOAT8OID:
			return GX_FLOAT8;
		case DATEOID:
			return GX_DATE;
		case CHAROID:
			return GX_CHAR;
		case BPCHAROID:
			return GX_CHAR;		/* bpchar(1) only: checked by the planner hook */
		default:
			return 0;
	}
}

/* ------------------------------------------------ K0: heap page loader
 * One heapgetpage() per block (pins the buffer, runs HeapTupleSatisfiesMVCC once
 * per tuple, fills rs_vistuples[]); the raw page bytes and the visible line
 * pointers go to the device in batches, the deform runs there. */
#define LOAD_BATCH_PAGES 4096	/* 32 MB of pages per gx_table_append_heap_pages() */

static void
gpuexec_load_relation(Relation rel, EState *estate, GpuRelInfo *info, gx_table **out)
{
	TupleDesc	desc = RelationGetDescr(rel);
	HeapScanDesc scan = heap_beginscan(rel, estate->es_snapshot, 0, NULL);
	BlockNumber nblocks = scan->rs_nblocks;
	gx_heap_desc hd;
	char	   *pages;
	uint16	   *vis;
	int32	   *cnt;
	int64		est_rows = (int64) nblocks * MaxHeapTuplesPerPage;
	BlockNumber blk;
	int			nbatch = 0;
	int			i;

	memset(&hd, 0, sizeof(hd));
	hd.natts = desc->natts;
	hd.ncols = info->ncols;
	if (desc->natts > 64)
		ereport(ERROR, (errmsg("gpuexec: relation has more than 64 attributes")));
	for (i = 0; i < desc->natts; i++)
	{
		Form_pg_attribute att = TupleDescAttr(desc, i);

		hd.att_len[i] = att->attlen;
		hd.att_align[i] = att->attalign == 'd' ? 8 : att->attalign == 'i' ? 4 : att->attalign == 's' ? 2 : 1;
	}
	for (i = 0; i < info->ncols; i++)
		hd.attnums[i] = info->attnums[i];

	GX_CHECK(gx_table_create(backend_ctx, info->ncols, info->types, est_rows > 0 ? est_rows : 1, out));
	/* staging lives in pinned memory so the copy is one DMA */
	GX_CHECK(gx_host_alloc(backend_ctx, (size_t) LOAD_BATCH_PAGES * BLCKSZ, (void **) &pages));
	vis = (uint16 *) palloc(sizeof(uint16) * LOAD_BATCH_PAGES * MaxHeapTuplesPerPage);
	cnt = (int32 *) palloc(sizeof(int32) * LOAD_BATCH_PAGES);

	for (blk = 0; blk < nblocks; blk++)
	{
		CHECK_FOR_INTERRUPTS();
		heapgetpage(scan, blk);
		LockBuffer(scan->rs_cbuf, BUFFER_LOCK_SHARE);
		memcpy(pages + (size_t) nbatch * BLCKSZ, BufferGetPage(scan->rs_cbuf), BLCKSZ);
		LockBuffer(scan->rs_cbuf, BUFFER_LOCK_UNLOCK);
		cnt[nbatch] = scan->rs_ntuples;
		memcpy(vis + (size_t) nbatch * MaxHeapTuplesPerPage, scan->rs_vistuples, sizeof(uint16) * scan->rs_ntuples);
		if (++nbatch == LOAD_BATCH_PAGES || blk + 1 == nblocks)
		{
			int			st = gx_table_append_heap_pages(*out, pages, nbatch, &hd, vis, cnt, MaxHeapTuplesPerPage);

			if (st != GX_OK)
			{
				gx_host_free(backend_ctx, pages);
				heap_endscan(scan);
				GX_CHECK(st);
			}
			nbatch = 0;
		}
	}
	gx_host_free(backend_ctx, pages);
	pfree(vis);
	pfree(cnt);
	heap_endscan(scan);
}

/* ------------------------------------------- custom_private (de)serialisation
 * A flat list of Integer / Float Value nodes — the only thing nodeToString()
 * is guaranteed to carry to the datanodes (nodes/outfuncs.c:1288-1310). */
static List *
put_int(List *l, int64 v)
{
	return lappend(l, makeInteger((int) v));
}
static List *
put_double(List *l, double v)
{
	char		buf[64];

	snprintf(buf, sizeof(buf), "%.17g", v);
	return lappend(l, makeFloat(pstrdup(buf)));
}
static int
get_int(ListCell **lc)
{
	int			v = intVal(lfirst(*lc));

	*lc = lnext(*lc);
	return v;
}
static double
get_double(ListCell **lc)
{
	double		v = floatVal(lfirst(*lc));

	*lc = lnext(*lc);
	return v;
}

static List *
serialise_rel(List *l, const GpuRelInfo *r)
{
	int			i;

	l = put_int(l, r->rti);
	l = put_int(l, r->ncols);
	for (i = 0; i < r->ncols; i++)
	{
		l = put_int(l, r->attnums[i]);
		l = put_int(l, r->types[i]);
	}
	return l;
}
static void
deserialise_rel(ListCell **lc, GpuRelInfo *r)
{
	int			i;

	r->rti = get_int(lc);
	r->ncols = get_int(lc);
	for (i = 0; i < r->ncols; i++)
	{
		r->attnums[i] = get_int(lc);
		r->types[i] = get_int(lc);
	}
}

static List *
gpuexec_serialise(const GpuExecState *st)
{
	List	   *l = NIL;
	const gx_agg_plan *p = &st->plan;
	int			i,
				j;

	l = put_int(l, GX_ABI_VERSION);
	l = put_int(l, st->has_join);
	l = serialise_rel(l, &st->outer);
	if (st->has_join)
	{
		l = serialise_rel(l, &st->inner);
		l = put_int(l, st->inner_key_col);
		l = put_int(l, st->n_payload);
		for (i = 0; i < st->n_payload; i++)
			l = put_int(l, st->payload_cols[i]);
		l = put_int(l, st->inner_unique);
	}
	l = put_int(l, p->n_preds);
	l = put_int(l, p->outer_key_col);
	for (i = 0; i < p->n_preds; i++)
	{
		l = put_int(l, p->preds[i].col);
		l = put_int(l, p->preds[i].op);
		l = put_double(l, (double) p->preds[i].ival);
		l = put_double(l, p->preds[i].fval);
	}
	l = put_int(l, p->n_group_cols);
	for (i = 0; i < p->n_group_cols; i++)
	{
		l = put_int(l, p->group_cols[i].side);
		l = put_int(l, p->group_cols[i].col);
	}
	l = put_int(l, p->n_aggs);
	for (i = 0; i < p->n_aggs; i++)
	{
		l = put_int(l, p->aggs[i].fn);
		l = put_int(l, p->aggs[i].arg.nops);
		for (j = 0; j < p->aggs[i].arg.nops; j++)
		{
			l = put_int(l, p->aggs[i].arg.ops[j].op);
			l = put_int(l, p->aggs[i].arg.ops[j].col);
			l = put_double(l, p->aggs[i].arg.ops[j].k);
		}
	}
	l = put_double(l, (double) p->est_groups);
	return l;
}

static void
gpuexec_deserialise(List *priv, GpuExecState *st)
{
	ListCell   *lc = list_head(priv);
	gx_agg_plan *p = &st->plan;
	int			i,
				j;

	if (get_int(&lc) != GX_ABI_VERSION)
		ereport(ERROR, (errmsg("gpuexec: coordinator and datanode run different provider versions")));
	memset(p, 0, sizeof(*p));
	st->has_join = get_int(&lc);
	deserialise_rel(&lc, &st->outer);
	if (st->has_join)
	{
		deserialise_rel(&lc, &st->inner);
		st->inner_key_col = get_int(&lc);
		st->n_payload = get_int(&lc);
		for (i = 0; i < st->n_payload; i++)
			st->payload_cols[i] = get_int(&lc);
		st->inner_unique = get_int(&lc);
	}
	p->n_preds = get_int(&lc);
	p->outer_key_col = get_int(&lc);
	for (i = 0; i < p->n_preds; i++)
	{
		p->preds[i].col = get_int(&lc);
		p->preds[i].op = get_int(&lc);
		p->preds[i].ival = (int64) get_double(&lc);
		p->preds[i].fval = get_double(&lc);
	}
	p->n_group_cols = get_int(&lc);
	for (i = 0; i < p->n_group_cols; i++)
	{
		p->group_cols[i].side = get_int(&lc);
		p->group_cols[i].col = get_int(&lc);
	}
	p->n_aggs = get_int(&lc);
	for (i = 0; i < p->n_aggs; i++)
	{
		p->aggs[i].fn = get_int(&lc);
		p->aggs[i].arg.nops = get_int(&lc);
		for (j = 0; j < p->aggs[i].arg.nops; j++)
		{
			p->aggs[i].arg.ops[j].op = get_int(&lc);
			p->aggs[i].arg.ops[j].col = get_int(&lc);
			p->aggs[i].arg.ops[j].k = get_double(&lc);
		}
	}
	p->est_groups = (int64) get_double(&lc);
}

/* ------------------------------------------------------ executor methods */
static void gpuexec_begin(CustomScanState *node, EState *estate, int eflags);
static TupleTableSlot *gpuexec_exec(CustomScanState *node);
static void gpuexec_end(CustomScanState *node);
static void gpuexec_rescan(CustomScanState *node);
static void gpuexec_explain(CustomScanState *node, List *ancestors, ExplainState *es);

static const CustomExecMethods gpuexec_exec_methods = {
	.CustomName = GPUEXEC_NAME,
	.BeginCustomScan = gpuexec_begin,
	.ExecCustomScan = gpuexec_exec,
	.EndCustomScan = gpuexec_end,
	.ReScanCustomScan = gpuexec_rescan,
	.ExplainCustomScan = gpuexec_explain,
};

static Node *
gpuexec_create_state(CustomScan *cscan)
{
	GpuExecState *st = (GpuExecState *) palloc0(sizeof(GpuExecState));

	NodeSetTag(st, T_CustomScanState);
	st->css.methods = &gpuexec_exec_methods;
	gpuexec_deserialise(cscan->custom_private, st);
	return (Node *) st;
}

static const CustomScanMethods gpuexec_scan_methods = {
	.CustomName = GPUEXEC_NAME,
	.CreateCustomScanState = gpuexec_create_state,
};

static void
gpuexec_begin(CustomScanState *node, EState *estate, int eflags)
{
	GpuExecState *st = (GpuExecState *) node;
	GpuHandles *h;

	if (eflags & EXEC_FLAG_EXPLAIN_ONLY)
		return;
	gpuexec_ensure_context();
	/* scanrelid == 0: ExecInitCustomScan built the scan slot from custom_scan_tlist
	 * (nodeCustom.c:81-94); we open the base relations ourselves */
	st->outer_rel = ExecOpenScanRelation(estate, st->outer.rti, eflags);
	if (st->has_join)
		st->inner_rel = ExecOpenScanRelation(estate, st->inner.rti, eflags);
	h = (GpuHandles *) malloc(sizeof(GpuHandles));
	h->owner = CurrentResourceOwner;
	h->state = st;
	h->next = live_handles;
	live_handles = h;
}

static void
gpuexec_run(GpuExecState *st, EState *estate)
{
	int64		n;
	int			ng = st->plan.n_group_cols,
				na = st->plan.n_aggs;
	instr_time	t0,
				t1;

	INSTR_TIME_SET_CURRENT(t0);
	if (st->has_join)
		gpuexec_load_relation(st->inner_rel, estate, &st->inner, &st->inner_tab);
	gpuexec_load_relation(st->outer_rel, estate, &st->outer, &st->outer_tab);
	INSTR_TIME_SET_CURRENT(t1);
	INSTR_TIME_SUBTRACT(t1, t0);
	st->load_ms = INSTR_TIME_GET_MILLISEC(t1);

	INSTR_TIME_SET_CURRENT(t0);
	if (st->has_join)
		GX_CHECK(gx_hash_build(backend_ctx, st->inner_tab, st->inner_key_col, 0, NULL,
							   st->n_payload, st->payload_cols, st->inner_unique, &st->hash));
	CHECK_FOR_INTERRUPTS();
	GX_CHECK(gx_hash_agg(backend_ctx, st->outer_tab, st->hash, &st->plan, &st->result));
	/* Partial -> Distribute -> Finalize happens in the planner's own RemoteSubplan
	 * above us unless all datanodes share one box; then gx_result_combine() does it
	 * over NVLink (no-op without a communicator). */
	GX_CHECK(gx_result_combine(backend_ctx, st->result));
	n = gx_result_ngroups(st->result);
	st->keys = (int64 *) palloc(sizeof(int64) * Max(n * ng, 1));
	st->aggs = (double *) palloc(sizeof(double) * Max(n * na, 1));
	st->nulls = (uint8 *) palloc(Max(n * (ng + na), 1));
	GX_CHECK(gx_result_fetch(st->result, n, st->keys, st->aggs, st->nulls));
	st->ngroups = n;
	st->next = 0;
	INSTR_TIME_SET_CURRENT(t1);
	INSTR_TIME_SUBTRACT(t1, t0);
	st->exec_ms = INSTR_TIME_GET_MILLISEC(t1);
	/* the device copies are no longer needed */
	gpuexec_release_state(st);
	st->done_exec = true;
}

/* Returns one group per call as a virtual tuple: the custom_scan_tlist is
 * (group columns..., aggregates...) in plan order. */
static TupleTableSlot *
gpuexec_exec(CustomScanState *node)
{
	GpuExecState *st = (GpuExecState *) node;
	TupleTableSlot *slot = node->ss.ss_ScanTupleSlot;
	TupleDesc	desc = slot->tts_tupleDescriptor;
	int			ng = st->plan.n_group_cols,
				na = st->plan.n_aggs;
	int			i;
	int64		g;

	if (!st->done_exec)
		gpuexec_run(st, node->ss.ps.state);
	ExecClearTuple(slot);
	if (st->next >= st->ngroups)
		return slot;			/* empty slot = end of scan */
	g = st->next++;
	for (i = 0; i < ng + na; i++)
	{
		bool		isnull = st->nulls[g * (ng + na) + i] != 0;
		Oid			typid = TupleDescAttr(desc, i)->atttypid;

		slot->tts_isnull[i] = isnull;
		if (isnull)
		{
			slot->tts_values[i] = (Datum) 0;
			continue;
		}
		if (i < ng)
		{
			int64		v = st->keys[g * ng + i];

			switch (typid)
			{
				case FLOAT8OID:
					{
						double		d;

						memcpy(&d, &v, sizeof(d));
						slot->tts_values[i] = Float8GetDatum(d);
						break;
					}
				case INT8OID:
					slot->tts_values[i] = Int64GetDatum(v);
					break;
				case CHAROID:
					slot->tts_values[i] = CharGetDatum((char) v);
					break;
				default:
					slot->tts_values[i] = Int32GetDatum((int32) v);
					break;
			}
		}
		else
		{
			double		d = st->aggs[g * na + (i - ng)];

			if (typid == INT8OID)
			{
				int64		v;

				memcpy(&v, &d, sizeof(v));
				slot->tts_values[i] = Int64GetDatum(v);
			}
			else
				slot->tts_values[i] = Float8GetDatum(d);
		}
	}
	return ExecStoreVirtualTuple(slot);
}

static void
gpuexec_end(CustomScanState *node)
{
	GpuExecState *st = (GpuExecState *) node;
	GpuHandles **p = &live_handles;

	gpuexec_release_state(st);
	while (*p)
	{
		if ((*p)->state == st)
		{
			GpuHandles *h = *p;

			*p = h->next;
			free(h);
		}
		else
			p = &(*p)->next;
	}
	if (st->outer_rel)
		ExecCloseScanRelation(st->outer_rel);
	if (st->inner_rel)
		ExecCloseScanRelation(st->inner_rel);
}

static void
gpuexec_rescan(CustomScanState *node)
{
	GpuExecState *st = (GpuExecState *) node;

	/* the result does not depend on parameters: replay it */
	st->next = 0;
}

static void
gpuexec_explain(CustomScanState *node, List *ancestors, ExplainState *es)
{
	GpuExecState *st = (GpuExecState *) node;

	ExplainPropertyText("GPU Strategy", st->has_join ? "hash build + fused probe/aggregate" : "scan + hash aggregate", es);
	if (es->analyze && st->done_exec)
	{
		ExplainPropertyFloat("GPU Load", "ms", st->load_ms, 3, es);
		ExplainPropertyFloat("GPU Exec", "ms", st->exec_ms, 3, es);
		ExplainPropertyInteger("GPU Groups", NULL, st->ngroups, es);
	}
}

/* ------------------------------------------------------- planner side */
static Plan *gpuexec_plan_path(PlannerInfo *root, RelOptInfo *rel, CustomPath *best_path,
				  List *tlist, List *clauses, List *custom_plans);

static const CustomPathMethods gpuexec_path_methods = {
	.CustomName = GPUEXEC_NAME,
	.PlanCustomPath = gpuexec_plan_path,
};

/* PlanCustomPath: the CustomPath carries the serialised descriptor; the scan
 * tuple is described by custom_scan_tlist = the upper rel's target list
 * (scanrelid = 0: nodeCustom.c:81-94 requires it). */
static Plan *
gpuexec_plan_path(PlannerInfo *root, RelOptInfo *rel, CustomPath *best_path,
				  List *tlist, List *clauses, List *custom_plans)
{
	CustomScan *cscan = makeNode(CustomScan);

	cscan->scan.plan.targetlist = tlist;
	cscan->scan.plan.qual = NIL;	/* HAVING is declined by the hook */
	cscan->scan.scanrelid = 0;
	cscan->flags = best_path->flags;
	cscan->custom_plans = NIL;
	cscan->custom_exprs = NIL;
	cscan->custom_private = best_path->custom_private;
	cscan->custom_scan_tlist = tlist;
	cscan->custom_relids = NULL;
	cscan->methods = &gpuexec_scan_methods;
	return &cscan->scan.plan;
}

/* ---- plan-shape analysis -------------------------------------------------
 * Accepts:   Agg(HASHED/PLAIN) over  SeqScan(R)                       (configs 1, 2)
 *            Agg over HashJoin[INNER, one int4/int8 equi-key](SeqScan(R), SeqScan(S))   (config 3)
 * with Var group keys, no HAVING / grouping sets / DISTINCT / ORDER BY aggregates /
 * FILTER, and aggregates count(*), count(x), sum/avg/min/max(float8 expr),
 * sum(int4).  Anything else returns false and the CPU paths stay untouched —
 * the same "eligibility test" idea as jit_compile_hashjoin (jit/jit.c:253-303). */
#define AGG_COUNT_STAR 2803
#define AGG_COUNT_ANY  2147
#define AGG_SUM_F8     2111
#define AGG_AVG_F8     2105
#define AGG_MAX_F8     2120
#define AGG_MIN_F8     2136
#define AGG_SUM_I4     2108

/* column number of (rti, attno) in the descriptor, appended on first use */
static int
rel_column(GpuRelInfo *r, Index rti, AttrNumber attno, Oid typid, int32 typmod)
{
	int32		gxt = gpuexec_type_of(typid);
	int			i;

	if (r->rti != rti || attno <= 0 || gxt == 0)
		return -1;
	if (typid == BPCHAROID && typmod != VARHDRSZ + 1)
		return -1;				/* only bpchar(1) is staged as one byte */
	for (i = 0; i < r->ncols; i++)
		if (r->attnums[i] == attno - 1)
			return i;
	if (r->ncols >= GX_MAX_COLS)
		return -1;
	r->attnums[r->ncols] = attno - 1;
	r->types[r->ncols] = gxt;
	return r->ncols++;
}

/* float8 expression -> postfix program over outer columns */
static bool
match_f8_expr(Node *n, GpuRelInfo *outer, gx_expr *e)
{
	if (n == NULL || e->nops >= GX_MAX_EXPR_OPS)
		return false;
	if (IsA(n, RelabelType))
		return match_f8_expr((Node *) ((RelabelType *) n)->arg, outer, e);
	if (IsA(n, Var))
	{
		Var		   *v = (Var *) n;
		int			c = rel_column(outer, v->varno, v->varattno, v->vartype, v->vartypmod);

		if (c < 0 || v->varlevelsup != 0 || (v->vartype != FLOAT8OID && v->vartype != INT4OID && v->vartype != INT8OID))
			return false;
		e->ops[e->nops].op = GX_OP_COL;
		e->ops[e->nops].col = c;
		e->nops++;
		return true;
	}
	if (IsA(n, Const))
	{
		Const	   *c = (Const *) n;

		if (c->constisnull || c->consttype != FLOAT8OID)
			return false;
		e->ops[e->nops].op = GX_OP_CONST;
		e->ops[e->nops].k = DatumGetFloat8(c->constvalue);
		e->nops++;
		return true;
	}
	if (IsA(n, OpExpr))
	{
		OpExpr	   *o = (OpExpr *) n;
		int			op;

		if (list_length(o->args) != 2)
			return false;
		switch (o->opfuncid)
		{
			case F_FLOAT8PL:
				op = GX_OP_ADD;
				break;
			case F_FLOAT8MI:
				op = GX_OP_SUB;
				break;
			case F_FLOAT8MUL:
				op = GX_OP_MUL;
				break;
			default:
				return false;
		}
		if (!match_f8_expr(linitial(o->args), outer, e) || !match_f8_expr(lsecond(o->args), outer, e))
			return false;
		if (e->nops >= GX_MAX_EXPR_OPS)
			return false;
		e->ops[e->nops++].op = op;
		return true;
	}
	return false;
}

static bool
gpuexec_match_plan(PlannerInfo *root, RelOptInfo *input_rel, RelOptInfo *output_rel, GpuExecState *out,
				   double *est_rows, double *est_groups)
{
	Query	   *parse = root->parse;
	Path	   *in = input_rel->cheapest_total_path;
	gx_agg_plan *plan = &out->plan;
	ListCell   *lc;

	if (parse->groupingSets || parse->havingQual || parse->hasWindowFuncs || parse->hasTargetSRFs || in == NULL)
		return false;
	plan->outer_key_col = -1;

	if (in->pathtype == T_SeqScan && input_rel->reloptkind == RELOPT_BASEREL)
	{
		if (input_rel->baserestrictinfo != NIL)
			return false;		/* quals: a later version maps "Var op Const" onto gx_pred */
		out->outer.rti = input_rel->relid;
	}
	else if (IsA(in, HashPath))
	{
		HashPath   *hp = (HashPath *) in;
		Path	   *op = hp->jpath.outerjoinpath,
				   *ip = hp->jpath.innerjoinpath;
		RestrictInfo *ri;
		OpExpr	   *clause;
		Var		   *lv,
				   *rv;

		if (hp->jpath.jointype != JOIN_INNER || list_length(hp->path_hashclauses) != 1 || hp->jpath.joinrestrictinfo == NIL ||
			list_length(hp->jpath.joinrestrictinfo) != 1)
			return false;
		if (op->pathtype != T_SeqScan || ip->pathtype != T_SeqScan ||
			op->parent->baserestrictinfo != NIL || ip->parent->baserestrictinfo != NIL)
			return false;
		ri = (RestrictInfo *) linitial(hp->path_hashclauses);
		clause = (OpExpr *) ri->clause;
		if (!IsA(clause, OpExpr) || list_length(clause->args) != 2 || !IsA(linitial(clause->args), Var) || !IsA(lsecond(clause->args), Var))
			return false;
		if (clause->opfuncid != F_INT4EQ && clause->opfuncid != F_INT8EQ)
			return false;
		lv = (Var *) linitial(clause->args);
		rv = (Var *) lsecond(clause->args);
		if (lv->varno != op->parent->relid)
		{
			Var		   *t = lv;

			lv = rv;
			rv = t;
		}
		if (lv->varno != op->parent->relid || rv->varno != ip->parent->relid)
			return false;
		out->has_join = true;
		out->outer.rti = op->parent->relid;
		out->inner.rti = ip->parent->relid;
		plan->outer_key_col = rel_column(&out->outer, lv->varno, lv->varattno, lv->vartype, lv->vartypmod);
		out->inner_key_col = rel_column(&out->inner, rv->varno, rv->varattno, rv->vartype, rv->vartypmod);
		out->inner_unique = hp->jpath.inner_unique;
		if (plan->outer_key_col < 0 || out->inner_key_col < 0)
			return false;
	}
	else
		return false;

	/* GROUP BY: plain Vars of either side; inner-side keys ride in the join payload */
	foreach(lc, parse->groupClause)
	{
		SortGroupClause *sgc = (SortGroupClause *) lfirst(lc);
		TargetEntry *tle = get_sortgroupclause_tle(sgc, parse->targetList);
		Var		   *v = (Var *) tle->expr;
		int			c;

		if (!IsA(v, Var) || plan->n_group_cols >= GX_MAX_GROUP_COLS)
			return false;
		if (v->varno == out->outer.rti)
		{
			c = rel_column(&out->outer, v->varno, v->varattno, v->vartype, v->vartypmod);
			if (c < 0)
				return false;
			plan->group_cols[plan->n_group_cols].side = 0;
			plan->group_cols[plan->n_group_cols].col = c;
		}
		else if (out->has_join && v->varno == out->inner.rti && out->n_payload < GX_MAX_PAYLOAD)
		{
			c = rel_column(&out->inner, v->varno, v->varattno, v->vartype, v->vartypmod);
			if (c < 0)
				return false;
			out->payload_cols[out->n_payload] = c;
			plan->group_cols[plan->n_group_cols].side = 1;
			plan->group_cols[plan->n_group_cols].col = out->n_payload++;
		}
		else
			return false;
		plan->n_group_cols++;
	}

	/* aggregates of the target list, in order (they follow the group columns in the scan tuple) */
	foreach(lc, parse->targetList)
	{
		TargetEntry *tle = (TargetEntry *) lfirst(lc);
		Aggref	   *a = (Aggref *) tle->expr;
		gx_agg	   *g;

		if (IsA(a, Var))
			continue;			/* a group column */
		if (!IsA(a, Aggref) || a->aggdistinct || a->aggorder || a->aggfilter || a->aggdirectargs || plan->n_aggs >= GX_MAX_AGGS)
			return false;
		g = &plan->aggs[plan->n_aggs];
		switch (a->aggfnoid)
		{
			case AGG_COUNT_STAR:
				g->fn = GX_AGG_COUNT_STAR;
				break;
			case AGG_SUM_F8:
				g->fn = GX_AGG_SUM_F8;
				break;
			case AGG_AVG_F8:
				g->fn = GX_AGG_AVG_F8;
				break;
			case AGG_MIN_F8:
				g->fn = GX_AGG_MIN_F8;
				break;
			case AGG_MAX_F8:
				g->fn = GX_AGG_MAX_F8;
				break;
			case AGG_SUM_I4:
				g->fn = GX_AGG_SUM_I4;
				break;
			case AGG_COUNT_ANY:
				g->fn = GX_AGG_COUNT;
				break;
			default:
				return false;
		}
		if (g->fn != GX_AGG_COUNT_STAR)
		{
			TargetEntry *arg;

			if (list_length(a->args) != 1)
				return false;
			arg = (TargetEntry *) linitial(a->args);
			if (!match_f8_expr((Node *) arg->expr, &out->outer, &g->arg))
				return false;
			if ((g->fn == GX_AGG_SUM_I4 || g->fn == GX_AGG_COUNT) && g->arg.nops != 1)
				return false;
		}
		plan->n_aggs++;
	}
	if (plan->n_aggs == 0 && plan->n_group_cols == 0)
		return false;
	*est_rows = in->rows;
	*est_groups = output_rel->rows > 0 ? output_rel->rows : 1;
	plan->est_groups = (int64) *est_groups;
	return true;
}

static void
gpuexec_upper_paths_hook(PlannerInfo *root, UpperRelationKind stage, RelOptInfo *input_rel, RelOptInfo *output_rel)
{
	GpuExecState desc;
	double		rows,
				groups;

	if (prev_upper_paths_hook)
		prev_upper_paths_hook(root, stage, input_rel, output_rel);
	if (!gpuexec_enabled || stage != UPPERREL_GROUP_AGG)
		return;
	memset(&desc, 0, sizeof(desc));
	if (!gpuexec_match_plan(root, input_rel, output_rel, &desc, &rows, &groups))
		return;					/* decline: the CPU paths stay as they are */
	{
		CustomPath *cpath = makeNode(CustomPath);
		Path	   *cheapest_in = input_rel->cheapest_total_path;

		cpath->path.pathtype = T_CustomScan;
		cpath->path.parent = output_rel;
		cpath->path.pathtarget = output_rel->reltarget;
		cpath->path.param_info = NULL;
		cpath->path.parallel_aware = false;
		cpath->path.parallel_safe = false;	/* the GPU replaces intra-node parallelism */
		cpath->path.rows = groups;
		/* staging dominates: charge the sequential page reads, nothing per tuple */
		cpath->path.startup_cost = cheapest_in->total_cost * 0.25;
		cpath->path.total_cost = cpath->path.startup_cost + groups * 0.01;
		/* the aggregate runs where the data lives: keep the input's distribution, the XL
		 * planner then adds the RemoteSubplan above us exactly as for a CPU HashAggregate
		 * (optimizer/util/pathnode.c:4575) */
		cpath->path.distribution = cheapest_in->distribution;
		cpath->flags = 0;
		cpath->custom_paths = NIL;
		cpath->custom_private = gpuexec_serialise(&desc);
		cpath->methods = &gpuexec_path_methods;
		add_path(output_rel, &cpath->path);
	}
}

void
_PG_init(void)
{
	DefineCustomBoolVariable("gpuexec.enabled", "Offer GPU paths for scan/join/aggregate sub-plans.", NULL,
							 &gpuexec_enabled, true, PGC_USERSET, 0, NULL, NULL, NULL);
	DefineCustomIntVariable("gpuexec.device", "CUDA device ordinal used by this datanode.", NULL,
							&gpuexec_device, 0, 0, 63, PGC_BACKEND, 0, NULL, NULL, NULL);
	RegisterCustomScanMethods(&gpuexec_scan_methods);
	RegisterResourceReleaseCallback(gpuexec_resowner_callback, NULL);
	prev_upper_paths_hook = create_upper_paths_hook;
	create_upper_paths_hook = gpuexec_upper_paths_hook;
}
