/*
 * harness_main.c — a stub-linked fake backend that EXECUTES the CustomScan provider.
 *
 * TEST INFRASTRUCTURE (tests/test_provider_harness.py).  The provider's source is included
 * verbatim (so its static functions are reachable); everything a live backend would supply
 * is faked here, the same way the reference's own unit tests link backend code against
 * stubs (src/backend/unittest/backend/stub/):
 *   - memory: palloc -> malloc; ereport(ERROR) prints and exits 3 (no PG_TRY is needed: the
 *     provider keeps no host resources across a longjmp);
 *   - relations: heap-page images read from a case file (built by the oracle's page writer,
 *     whose bytes are pinned against the reference's heaptuple.o / bufpage.o);
 *     heap_beginscan / heapgetpage / heap_endscan walk them with the REFERENCE's page and
 *     tuple macros (bufpage.h, itemid.h, htup_details.h) and fill rs_vistuples exactly as
 *     heapgetpage() does for an all-committed snapshot; BufferGetPage() resolves through the
 *     real macro (local-buffer branch of BufferGetBlock);
 *   - executor: what ExecInitCustomScan does around the provider (nodeCustom.c:32-115) is
 *     restated in main(): CreateCustomScanState, a scan slot with the scan tuple's
 *     descriptor, BeginCustomScan, ExecCustomScan until TupIsNull, EndCustomScan;
 *   - lists and Value nodes are the reference's own object code (nodes/list.c, nodes/value.c,
 *     compiled from where they lie).
 * libgpuexec.so is the real thing: the harness needs a GPU.
 *
 * usage: gpuexec_harness <case file>     -> one text line per result row on stdout
 */
#include "../gpuexec_provider.c"

#include "storage/itemid.h"
#include "storage/bufpage.h"

#include <stdio.h>
#include <stdlib.h>

#ifndef FLOAT8ARRAYOID
#define FLOAT8ARRAYOID 1022		/* _float8, catalog/pg_type.h */
#endif

/* ------------------------------------------------------------ fake backend */
void *palloc_internal(Size size, const char *file, int line) { return malloc(size ? size : 1); }
void *palloc0_internal(Size size, const char *file, int line) { return calloc(1, size ? size : 1); }
void *palloc_extended_internal(Size size, int flags, const char *file, int line) { return calloc(1, size ? size : 1); }
void *repallocInternal(void *p, Size size, const char *file, int line) { return realloc(p, size); }
void pfree(void *p) { free(p); }
void *MemoryContextAllocInternal(MemoryContext c, Size size, const char *file, int line) { return malloc(size ? size : 1); }
void *MemoryContextAllocZeroInternal(MemoryContext c, Size size, const char *file, int line) { return calloc(1, size ? size : 1); }
void *MemoryContextAllocZeroAlignedIternal(MemoryContext c, Size size, const char *file, int line) { return calloc(1, size ? size : 1); }
char *pstrdup(const char *in) { return strdup(in); }
void MemoryContextReset(MemoryContext c) { }
__thread MemoryContext CurrentMemoryContext = NULL;
__thread ResourceOwner CurrentResourceOwner = NULL;
volatile bool InterruptPending = false;
void ProcessInterrupts(void) { }
void pg_qsort(void *base, size_t nel, size_t elsize, int (*cmp) (const void *, const void *)) { qsort(base, nel, elsize, cmp); }
__thread bool am_sub_thread = false;
void CheckSubThreadContextInternal(MemoryContext context, const char *func) { }
bool enable_resource_queue = false;
void ResQUsageBackoffPriority(void) { }
char *BufferBlocks = NULL;
Block *LocalBufferBlockPointers = NULL;
bool g_enable_cls = false, g_enable_data_mask = false, enable_fga = false;
create_upper_paths_hook_type create_upper_paths_hook = NULL;

static char last_msg[1024];
bool errstart(int elevel, const char *filename, int lineno, const char *funcname, const char *domain)
{ return elevel >= ERROR; }
void errfinish(int dummy,...) { fprintf(stderr, "harness: ereport(ERROR): %s\n", last_msg); exit(3); }
int errmsg(const char *fmt,...) { va_list ap; va_start(ap, fmt); vsnprintf(last_msg, sizeof(last_msg), fmt, ap); va_end(ap); return 0; }
int errmsg_internal(const char *fmt,...) { return 0; }
int errcode(int sqlerrcode) { return 0; }
int errdetail(const char *fmt,...) { return 0; }
void elog_start(const char *filename, int lineno, const char *funcname) { }
void elog_finish(int elevel, const char *fmt,...) { if (elevel >= ERROR) { fprintf(stderr, "harness: elog(ERROR): %s\n", fmt); exit(3); } }
void ExceptionalCondition(const char *a, const char *b, const char *c, int d) { fprintf(stderr, "harness: Assert(%s) at %s:%d\n", a, c, d); abort(); }

/* planner-side entry points: linked, never reached (the harness hands the provider a finished descriptor) */
#define NOT_REACHED(name) do { fprintf(stderr, "harness: %s() is planner-side and must not be reached\n", name); abort(); } while (0)
void add_path(RelOptInfo *r, Path *p) { NOT_REACHED("add_path"); }
AggPath *create_agg_path(PlannerInfo *root, RelOptInfo *rel, Path *subpath, PathTarget *target, AggStrategy s, AggSplit sp, List *g, List *q,
						 const AggClauseCosts *c, double n) { NOT_REACHED("create_agg_path"); return NULL; }
PathTarget *create_empty_pathtarget(void) { NOT_REACHED("create_empty_pathtarget"); return NULL; }
void add_column_to_pathtarget(PathTarget *t, Expr *e, Index r) { NOT_REACHED("add_column_to_pathtarget"); }
Path *create_redistribute_grouping_path(PlannerInfo *root, Query *parse, Path *path) { NOT_REACHED("create_redistribute_grouping_path"); return NULL; }
void get_agg_clause_costs(PlannerInfo *root, Node *clause, AggSplit s, AggClauseCosts *c) { NOT_REACHED("get_agg_clause_costs"); }
TargetEntry *get_sortgroupclause_tle(SortGroupClause *s, List *t) { NOT_REACHED("get_sortgroupclause_tle"); return NULL; }
void mark_partial_aggref(Aggref *a, AggSplit s) { NOT_REACHED("mark_partial_aggref"); }
TargetEntry *makeTargetEntry(Expr *e, AttrNumber r, char *n, bool j) { NOT_REACHED("makeTargetEntry"); return NULL; }
void *copyObjectImpl(const void *o) { NOT_REACHED("copyObject"); return NULL; }
bool equal(const void *a, const void *b) { NOT_REACHED("equal"); return false; }
bool cls_check_table_has_policy(Oid r) { return false; }
bool datamask_check_table_has_datamask(Oid r) { return false; }
bool get_audit_fga_quals(Oid rel, char *cmd, List *tl, List **out) { return false; }
void DefineCustomBoolVariable(const char *n, const char *s, const char *l, bool *v, bool b, GucContext c, int f, GucBoolCheckHook a, GucBoolAssignHook g, GucShowHook h) { }
void DefineCustomRealVariable(const char *n, const char *s, const char *l, double *v, double b, double mn, double mx, GucContext c, int f, GucRealCheckHook a, GucRealAssignHook g, GucShowHook h) { }
double		seq_page_cost = 1.0, cpu_tuple_cost = 0.01;		/* costsize.c:100-104 */
double path_count_datanodes(Path *p) { NOT_REACHED("path_count_datanodes"); return 1; }
void DefineCustomIntVariable(const char *n, const char *s, const char *l, int *v, int b, int mn, int mx, GucContext c, int f, GucIntCheckHook a, GucIntAssignHook g, GucShowHook h) { }
static const CustomScanMethods *registered_methods = NULL;
void RegisterCustomScanMethods(const CustomScanMethods *m) { registered_methods = m; }
static ResourceReleaseCallback release_cb = NULL;
void RegisterResourceReleaseCallback(ResourceReleaseCallback cb, void *arg) { release_cb = cb; }
void ExplainPropertyText(const char *q, const char *v, ExplainState *es) { fprintf(stderr, "explain: %s: %s\n", q, v); }
void ExplainPropertyFloat(const char *q, const char *u, double v, int nd, ExplainState *es) { fprintf(stderr, "explain: %s: %.3f %s\n", q, v, u ? u : ""); }
void ExplainPropertyInteger(const char *q, const char *u, int64 v, ExplainState *es) { fprintf(stderr, "explain: %s: %ld\n", q, (long) v); }

/* text/array constructors the executor side uses for by-reference datums */
text *cstring_to_text_with_len(const char *s, int len)
{
	text	   *t = (text *) malloc(len + VARHDRSZ);

	SET_VARSIZE(t, len + VARHDRSZ);
	memcpy(VARDATA(t), s, len);
	return t;
}
/* construct_array for a 1-D array of pass-by-value float8 without NULLs (utils/adt/arrayfuncs.c:3306) */
ArrayType *construct_array(Datum *elems, int nelems, Oid elmtype, int elmlen, bool elmbyval, char elmalign)
{
	Size		nbytes = ARR_OVERHEAD_NONULLS(1) + (Size) nelems * elmlen;
	ArrayType  *a = (ArrayType *) calloc(1, nbytes);
	int			i;

	if (elmtype != FLOAT8OID || elmlen != 8 || !elmbyval) { fprintf(stderr, "harness: construct_array: only float8[]\n"); abort(); }
	SET_VARSIZE(a, nbytes);
	a->ndim = 1; a->dataoffset = 0; a->elemtype = elmtype;
	ARR_DIMS(a)[0] = nelems; ARR_LBOUND(a)[0] = 1;
	for (i = 0; i < nelems; i++)
		memcpy(ARR_DATA_PTR(a) + (Size) i * 8, &elems[i], 8);
	return a;
}

/* ------------------------------------------------------------ fake relations */
typedef struct FakeRel
{
	RelationData rd;
	FormData_pg_class cls;
	int64		npages;
	char	   *pages;
	int			buf_base;		/* index of page 0 in LocalBufferBlockPointers */
} FakeRel;
static FakeRel fake_rels[2];
static int	n_fake_rels = 0;

Relation ExecOpenScanRelation(EState *estate, Index scanrelid, int eflags)
{
	if ((int) scanrelid < 1 || (int) scanrelid > n_fake_rels) { fprintf(stderr, "harness: no relation with rti %u\n", scanrelid); exit(3); }
	return &fake_rels[scanrelid - 1].rd;
}
void ExecCloseScanRelation(Relation r) { }

HeapScanDesc heap_beginscan(Relation relation, Snapshot snapshot, int nkeys, ScanKey key)
{
	HeapScanDesc scan = (HeapScanDesc) calloc(1, sizeof(HeapScanDescData));
	FakeRel    *fr = (FakeRel *) relation;

	scan->rs_rd = relation;
	scan->rs_snapshot = snapshot;
	scan->rs_nblocks = (BlockNumber) fr->npages;
	scan->rs_cbuf = InvalidBuffer;
	scan->rs_pageatatime = true;
	return scan;
}
void heap_endscan(HeapScanDesc scan) { free(scan); }
void LockBuffer(Buffer b, int mode) { }

/* heapgetpage (access/heap/heapam.c:388-513) for a snapshot that sees every committed insert and
 * no in-progress transaction: a tuple is visible iff xmin is committed and xmax is invalid */
void heapgetpage(HeapScanDesc scan, BlockNumber page)
{
	FakeRel    *fr = (FakeRel *) scan->rs_rd;
	Page		dp;
	int			lines,
				ntup = 0;
	OffsetNumber lineoff;

	scan->rs_cbuf = -(fr->buf_base + (int) page + 1);		/* a "local buffer": BufferGetBlock's negative branch */
	scan->rs_cblock = page;
	dp = BufferGetPage(scan->rs_cbuf);
	lines = PageGetMaxOffsetNumber(dp);
	for (lineoff = FirstOffsetNumber; lineoff <= lines; lineoff++)
	{
		ItemId		lpp = PageGetItemId(dp, lineoff);

		if (ItemIdIsNormal(lpp))
		{
			HeapTupleHeader th = (HeapTupleHeader) PageGetItem(dp, lpp);

			if ((th->t_infomask & HEAP_XMIN_COMMITTED) && (th->t_infomask & HEAP_XMAX_INVALID))
				scan->rs_vistuples[ntup++] = lineoff;
		}
	}
	scan->rs_ntuples = ntup;
}

/* ------------------------------------------------------------ slots (executor/execTuples.c:612,679) */
TupleTableSlot *ExecClearTuple(TupleTableSlot *slot)
{
	slot->tts_tuple = NULL;
	slot->tts_flags |= TTS_FLAG_EMPTY;
	slot->tts_nvalid = 0;
	return slot;
}
TupleTableSlot *ExecStoreVirtualTuple(TupleTableSlot *slot)
{
	slot->tts_flags &= ~TTS_FLAG_EMPTY;
	slot->tts_nvalid = slot->tts_tupleDescriptor->natts;
	return slot;
}

/* ------------------------------------------------------------ case file */
static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "harness: short case file\n"); exit(2); } }
static int32 rd32(FILE *f) { int32 v; rd(f, &v, 4); return v; }
static int64 rd64(FILE *f) { int64 v; rd(f, &v, 8); return v; }

static void read_relinfo(FILE *f, GpuRelInfo *r)
{
	int			i;

	r->rti = (Index) rd32(f);
	r->ncols = rd32(f);
	for (i = 0; i < r->ncols; i++) { r->attnums[i] = rd32(f); r->types[i] = rd32(f); }
}

/* read.c calls this for '{' nodes; custom_private only ever holds Value nodes */
Node *parseNodeString(void) { NOT_REACHED("parseNodeString"); return NULL; }

/* What nodeToString() writes for a List of Value nodes: _outList (nodes/outfuncs.c:443-477) around _outValue (:5103-5117:
 * T_Integer as %d, T_Float as its string, unquoted). */
static char *
out_value_list(List *l)
{
	size_t		cap = 64 + (size_t) list_length(l) * 40, n = 0;
	char	   *buf = (char *) malloc(cap);
	ListCell   *lc;

	buf[n++] = '(';
	foreach(lc, l)
	{
		Value	   *v = (Value *) lfirst(lc);

		if (IsA(v, Integer))
			n += snprintf(buf + n, cap - n, "%d", (int) intVal(v));
		else if (IsA(v, Float))
			n += snprintf(buf + n, cap - n, "%s", strVal(v));
		else
			NOT_REACHED("out_value_list: node type");
		if (lnext(lc))
			buf[n++] = ' ';
	}
	buf[n++] = ')';
	buf[n] = 0;
	return buf;
}

int main(int argc, char **argv)
{
	bool		ship_only = false;
	FILE	   *f;
	char		magic[4];
	int			nrels, i, r, nout, total_pages = 0;
	GpuExecState desc;
	Oid			out_types[GX_MAX_GROUP_COLS + GX_MAX_AGGS];
	CustomScan *cscan;
	CustomScanState *css;
	EState		estate;
	ExprContext econtext;
	TupleDesc	sdesc;
	TupleTableSlot *slot;
	int64		nrows = 0;

	if (argc >= 3 && strcmp(argv[1], "--ship-only") == 0) { ship_only = true; argv++; argc--; }
	if (argc < 2) { fprintf(stderr, "usage: %s [--ship-only] <case file>\n", argv[0]); return 2; }
	f = fopen(argv[1], "rb");
	if (!f) { perror(argv[1]); return 2; }
	rd(f, magic, 4);
	if (memcmp(magic, "GXH1", 4) != 0) { fprintf(stderr, "harness: bad magic\n"); return 2; }
	nrels = rd32(f);
	for (r = 0; r < nrels; r++)
	{
		FakeRel    *fr = &fake_rels[r];
		int			natts = rd32(f);
		TupleDesc	d = (TupleDesc) calloc(1, offsetof(struct tupleDesc, attrs) + natts * sizeof(FormData_pg_attribute));

		d->natts = natts;
		d->tdrefcount = -1;
		for (i = 0; i < natts; i++)
		{
			Form_pg_attribute a = TupleDescAttr(d, i);
			int32		len = rd32(f), align = rd32(f), typid = rd32(f), notnull = rd32(f);

			a->attlen = (int16) len;
			a->attalign = align == 8 ? 'd' : align == 4 ? 'i' : align == 2 ? 's' : 'c';
			a->atttypid = (Oid) typid;
			a->attnum = i + 1;
			a->attbyval = len > 0;
			a->attnotnull = notnull != 0;
			a->atttypmod = typid == BPCHAROID ? VARHDRSZ + 1 : -1;
			snprintf(NameStr(a->attname), NAMEDATALEN, "a%d", i + 1);
		}
		fr->npages = rd64(f);
		rd(f, &fr->cls.reltuples, 4);			/* float4, as in pg_class */
		fr->pages = (char *) malloc((size_t) fr->npages * BLCKSZ + 1);
		rd(f, fr->pages, (size_t) fr->npages * BLCKSZ);
		fr->rd.rd_att = d;
		fr->rd.rd_rel = &fr->cls;
		fr->buf_base = total_pages;
		total_pages += (int) fr->npages;
		n_fake_rels++;
	}
	LocalBufferBlockPointers = (Block *) calloc((size_t) total_pages + 1, sizeof(Block));
	for (r = 0; r < nrels; r++)
		for (i = 0; i < fake_rels[r].npages; i++)
			LocalBufferBlockPointers[fake_rels[r].buf_base + i] = (Block) (fake_rels[r].pages + (size_t) i * BLCKSZ);

	memset(&desc, 0, sizeof(desc));
	desc.has_join = rd32(f);
	read_relinfo(f, &desc.outer);
	if (desc.has_join)
	{
		read_relinfo(f, &desc.inner);
		desc.inner_key_col = rd32(f);
		desc.n_payload = rd32(f);
		for (i = 0; i < desc.n_payload; i++) desc.payload_cols[i] = rd32(f);
		desc.inner_unique = rd32(f);
		desc.n_inner_preds = rd32(f);
		rd(f, desc.inner_preds, sizeof(gx_pred) * desc.n_inner_preds);
	}
	desc.partial = rd32(f);
	rd(f, &desc.plan, sizeof(gx_agg_plan));
	nout = rd32(f);
	for (i = 0; i < nout; i++) out_types[i] = (Oid) rd32(f);
	fclose(f);

	if (ship_only)
	{
		/* CN -> DN plan shipping without a GPU: the descriptor as Value nodes, written the way nodeToString() writes them,
		 * read back by the reference's OWN nodes/read.c (stringToNode), deserialised and serialised again */
		List	   *priv = gpuexec_serialise(&desc);
		char	   *wire = out_value_list(priv);
		List	   *back = (List *) stringToNode(wire);
		GpuExecState got;
		char	   *again;

		if (!IsA(back, List) || list_length(back) != list_length(priv)) { fprintf(stderr, "harness: %d nodes sent, %d read back\n", list_length(priv), back ? list_length(back) : -1); return 3; }
		memset(&got, 0, sizeof(got));
		gpuexec_deserialise(back, &got);
		again = out_value_list(gpuexec_serialise(&got));
		if (strcmp(wire, again) != 0) { fprintf(stderr, "harness: descriptor changed in transit\n sent %s\n got  %s\n", wire, again); return 3; }
		for (i = 0; i < desc.plan.n_preds; i++)
			if (desc.plan.preds[i].ival != got.plan.preds[i].ival || memcmp(&desc.plan.preds[i].fval, &got.plan.preds[i].fval, 8) != 0) { fprintf(stderr, "harness: qual constant %d changed\n", i); return 3; }
		for (i = 0; i < desc.n_inner_preds; i++)
			if (desc.inner_preds[i].ival != got.inner_preds[i].ival || memcmp(&desc.inner_preds[i].fval, &got.inner_preds[i].fval, 8) != 0) { fprintf(stderr, "harness: inner qual constant %d changed\n", i); return 3; }
		if (desc.plan.est_groups != got.plan.est_groups) { fprintf(stderr, "harness: est_groups changed\n"); return 3; }
		printf("ship ok: %d nodes, %zu bytes\n%s\n", list_length(priv), strlen(wire), wire);
		return 0;
	}
	/* ---- what the planner + ExecInitCustomScan would have done */
	_PG_init();
	if (!registered_methods) { fprintf(stderr, "harness: _PG_init registered nothing\n"); return 3; }
	cscan = makeNode(CustomScan);
	cscan->methods = registered_methods;
	cscan->custom_private = gpuexec_serialise(&desc);	/* the descriptor travels as Value nodes, as it would CN -> DN */
	css = (CustomScanState *) cscan->methods->CreateCustomScanState(cscan);
	memset(&estate, 0, sizeof(estate));
	memset(&econtext, 0, sizeof(econtext));
	css->ss.ps.plan = &cscan->scan.plan;
	css->ss.ps.state = &estate;
	css->ss.ps.ps_ExprContext = &econtext;
	css->ss.ps.ps_ProjInfo = NULL;		/* target list == scan tuple: no projection step */
	sdesc = (TupleDesc) calloc(1, offsetof(struct tupleDesc, attrs) + nout * sizeof(FormData_pg_attribute));
	sdesc->natts = nout;
	for (i = 0; i < nout; i++) { TupleDescAttr(sdesc, i)->atttypid = out_types[i]; TupleDescAttr(sdesc, i)->attnum = i + 1; }
	slot = (TupleTableSlot *) calloc(1, sizeof(TupleTableSlot));
	slot->type = T_TupleTableSlot;
	slot->tts_flags = TTS_FLAG_EMPTY;
	slot->tts_tupleDescriptor = sdesc;
	slot->tts_values = (Datum *) calloc(nout, sizeof(Datum));
	slot->tts_isnull = (bool *) calloc(nout, sizeof(bool));
	css->ss.ss_ScanTupleSlot = slot;

	css->methods->BeginCustomScan(css, &estate, 0);
	for (;;)
	{
		TupleTableSlot *res = css->methods->ExecCustomScan(css);

		if (TupIsNull(res))
			break;
		for (i = 0; i < nout; i++)
		{
			if (i) putchar('\t');
			if (res->tts_isnull[i]) { fputs("\\N", stdout); continue; }
			switch (out_types[i])
			{
				case FLOAT8OID: printf("%.17g", DatumGetFloat8(res->tts_values[i])); break;
				case INT8OID: printf("%ld", (long) DatumGetInt64(res->tts_values[i])); break;
				case CHAROID: printf("%d", (int) (signed char) DatumGetChar(res->tts_values[i])); break;
				case BPCHAROID: printf("%d", (int) (signed char) VARDATA_ANY(DatumGetPointer(res->tts_values[i]))[0]); break;
				case FLOAT8ARRAYOID:
					{
						ArrayType  *a = (ArrayType *) DatumGetPointer(res->tts_values[i]);
						double	   *v = (double *) ARR_DATA_PTR(a);

						printf("{%.17g,%.17g,%.17g}", v[0], v[1], v[2]);
						break;
					}
				default: printf("%d", DatumGetInt32(res->tts_values[i])); break;
			}
		}
		putchar('\n');
		nrows++;
	}
	css->methods->ExplainCustomScan(css, NIL, &(ExplainState) {.analyze = true});
	css->methods->ReScanCustomScan(css);
	if (TupIsNull(css->methods->ExecCustomScan(css)) != (nrows == 0)) { fprintf(stderr, "harness: rescan did not replay the result\n"); return 3; }
	css->methods->EndCustomScan(css);
	if (live_handles != NULL) { fprintf(stderr, "harness: EndCustomScan left handles registered\n"); return 3; }
	gx_shutdown(backend_ctx);
	fprintf(stderr, "harness: %ld rows\n", (long) nrows);
	return 0;
}
