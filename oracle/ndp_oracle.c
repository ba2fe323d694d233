/* oracle/ndp_oracle.c -- TEST INFRASTRUCTURE ONLY (see ndp_oracle.h for the rules and the
 * parity pin).  Plain-C restatement of the reference's BFS granulation -> DFS hot path.
 * Each function names the reference lines it follows; the data structures are this file's own
 * (flat int32 triples, a ring of heap nodes, an explicit array stack).
 */
#include "ndp_oracle.h"

#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* parseDimacsString -- NDP-5_6_7.cpp:619-642.
 * Line based.  Lines that are empty or start with 'c' or 'p' are skipped (NDP:625).  Integers
 * are read with operator>> semantics until a read fails or yields 0 (NDP:630): a token that is
 * not an integer, or one that overflows int, ends the clause with the literals read so far.
 * Width 1 -> {0,0,x} (NDP:632-634); width 3 -> {x,y,z} (NDP:635-637); any other width is
 * silently dropped (NDP:639). */
static int read_int_like_istream(const char** pp, const char* end, int* out) {
    const char* p = *pp;
    while (p < end && isspace((unsigned char)*p)) ++p;
    if (p >= end) return 0;
    const char* q = p;
    if (*q == '+' || *q == '-') ++q;
    if (q >= end || !isdigit((unsigned char)*q)) return 0;
    long long v = 0;
    int overflow = 0;
    const char* d = q;
    while (d < end && isdigit((unsigned char)*d)) {
        if (!overflow) {
            v = v * 10 + (*d - '0');
            if (v > (long long)INT_MAX + 1) overflow = 1;
        }
        ++d;
    }
    if (*p == '-') v = -v;
    if (overflow || v > INT_MAX || v < INT_MIN) return 0; /* failbit: extraction fails */
    *out = (int)v;
    *pp = d;
    return 1;
}

size_t orc_parse_dimacs(const char* text, int32_t* out3, size_t cap_clauses) {
    size_t n = 0;
    const char* p = text;
    const char* end = text + strlen(text);
    while (p < end) {
        const char* eol = memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        if (eol > p && p[0] != 'c' && p[0] != 'p') {
            int lits[4];
            size_t nl = 0;
            const char* q = p;
            int v;
            while (read_int_like_istream(&q, eol, &v) && v != 0) {
                if (nl < 4) lits[nl] = v;
                ++nl;
            }
            if (nl == 1 || nl == 3) {
                if (out3 && n < cap_clauses) {
                    if (nl == 1) {
                        out3[3 * n] = 0; out3[3 * n + 1] = 0; out3[3 * n + 2] = lits[0];
                    } else {
                        out3[3 * n] = lits[0]; out3[3 * n + 1] = lits[1]; out3[3 * n + 2] = lits[2];
                    }
                }
                ++n;
            }
        }
        p = eol + 1;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* choice -- NDP-5_6_7.cpp:753-781.  First clause with exactly two zeros -> |its last non-zero
 * literal| (NDP:756-766); else first clause with exactly one zero -> |its LAST non-zero
 * literal| (NDP:767-777); else |A[0].l[0]| (NDP:778-779); 0 iff A is empty (NDP:780). */
static int abs_last_nonzero_if(const int32_t* cl, int want_zeros) {
    int zeros = 0, nonzero = 0;
    for (int j = 0; j < 3; ++j) {
        if (cl[j] == 0) ++zeros;
        else nonzero = cl[j];
    }
    if (zeros != want_zeros) return 0;
    return nonzero < 0 ? -nonzero : nonzero;
}

int orc_choice(const int32_t* c3, size_t n) {
    for (int want = 2; want >= 1; --want)
        for (size_t k = 0; k < n; ++k) {
            int v = abs_last_nonzero_if(c3 + 3 * k, want);
            if (v) return v;
        }
    if (n) return c3[0] < 0 ? -c3[0] : c3[0];
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* ResolutionStep (NDP:700-750) and ResolutionStepWithConflict (NDP:654-697) share one pass:
 * LA (variable i := true): a clause holding literal i is satisfied and dropped, literal -i
 * becomes 0; RA (i := false) is the mirror image.  Emission keeps input order.  The
 * with-conflict variant additionally flags a branch when an emitted clause is {0,0,0}
 * (NDP:685-686, 691-692). */
void orc_resolution_step(const int32_t* c3, size_t n, int i, int with_conflict, int32_t* out_la, size_t* n_la,
                         int32_t* out_ra, size_t* n_ra, int* conflict_la, int* conflict_ra) {
    size_t la = 0, ra = 0;
    int cla = 0, cra = 0;
    for (size_t k = 0; k < n; ++k) {
        const int32_t* cl = c3 + 3 * k;
        int sat_la = 0, sat_ra = 0;
        int32_t l_la[3], l_ra[3];
        for (int j = 0; j < 3; ++j) {
            int lit = cl[j];
            if (lit == i) { sat_la = 1; l_ra[j] = 0; l_la[j] = lit; }
            else if (lit == -i) { sat_ra = 1; l_la[j] = 0; l_ra[j] = lit; }
            else { l_la[j] = lit; l_ra[j] = lit; }
        }
        if (!sat_la) {
            if (!l_la[0] && !l_la[1] && !l_la[2]) cla = 1;
            memcpy(out_la + 3 * la, l_la, sizeof l_la);
            ++la;
        }
        if (!sat_ra) {
            if (!l_ra[0] && !l_ra[1] && !l_ra[2]) cra = 1;
            memcpy(out_ra + 3 * ra, l_ra, sizeof l_ra);
            ++ra;
        }
    }
    *n_la = la;
    *n_ra = ra;
    *conflict_la = with_conflict ? cla : -1;
    *conflict_ra = with_conflict ? cra : -1;
}

static int has_all_zero_clause(const int32_t* c3, size_t n) { /* the scans at NDP:905-907,916-918 */
    for (size_t k = 0; k < n; ++k)
        if (!c3[3 * k] && !c3[3 * k + 1] && !c3[3 * k + 2]) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t* cl;   /* 3*n ints, owned */
    size_t n;
    int32_t* ch;   /* choices, owned */
    size_t nch;
} node_t;

static node_t node_make(const int32_t* cl, size_t n, const int32_t* ch, size_t nch, int extra_lit, int has_extra) {
    node_t nd;
    nd.n = n;
    nd.cl = (int32_t*)malloc((n ? n : 1) * 3 * sizeof(int32_t));
    memcpy(nd.cl, cl, n * 3 * sizeof(int32_t));
    nd.nch = nch + (has_extra ? 1 : 0);
    nd.ch = (int32_t*)malloc((nd.nch ? nd.nch : 1) * sizeof(int32_t));
    if (nch) memcpy(nd.ch, ch, nch * sizeof(int32_t));
    if (has_extra) nd.ch[nch] = extra_lit;
    return nd;
}
static void node_free(node_t* nd) { free(nd->cl); free(nd->ch); nd->cl = NULL; nd->ch = NULL; }

struct orc_frontier {
    node_t* v;
    size_t head, tail, cap; /* live entries are v[head..tail) */
};
static void fq_push(orc_frontier* q, node_t nd) {
    if (q->tail == q->cap) {
        if (q->head > q->cap / 2) {
            memmove(q->v, q->v + q->head, (q->tail - q->head) * sizeof(node_t));
            q->tail -= q->head;
            q->head = 0;
        } else {
            q->cap = q->cap ? q->cap * 2 : 64;
            q->v = (node_t*)realloc(q->v, q->cap * sizeof(node_t));
        }
    }
    q->v[q->tail++] = nd;
}

/* Satisfy_iterative_BFS -- NDP-5_6_7.cpp:881-944. */
orc_frontier* orc_bfs(const int32_t* c3, size_t n, int max_iterations, int override_max_iterations, int max_queues,
                      int* iterations_out, int* task_count_out, size_t* initial_queue_size, uint64_t* clause_evals) {
    orc_frontier* q = (orc_frontier*)calloc(1, sizeof *q);
    fq_push(q, node_make(c3, n, NULL, 0, 0, 0));                               /* NDP:884-885 */
    int iterations = 0, task_count = 1;                                        /* NDP:887-888 */
    uint64_t evals = 0;
    size_t scratch_cap = n ? n : 1;
    int32_t* la = (int32_t*)malloc(scratch_cap * 3 * sizeof(int32_t));
    int32_t* ra = (int32_t*)malloc(scratch_cap * 3 * sizeof(int32_t));
    while (q->head < q->tail) {                                                /* NDP:893 */
        if (max_queues > 0 && q->tail - q->head >= (size_t)max_queues) break;  /* NDP:894-895 */
        if (max_iterations == -1 && !override_max_iterations && iterations >= max_iterations) break; /* NDP:896 */
        node_t cur = q->v[q->head++];                                          /* NDP:898-899 */
        int i = orc_choice(cur.cl, cur.n);                                     /* NDP:900 */
        if (i == 0) { node_free(&cur); continue; }                             /* NDP:901-902 */
        size_t nla, nra;
        int dummy1, dummy2;
        orc_resolution_step(cur.cl, cur.n, i, 0, la, &nla, ra, &nra, &dummy1, &dummy2); /* NDP:903 */
        evals += cur.n;
        if (nla && !has_all_zero_clause(la, nla)) {                            /* NDP:904-914 */
            fq_push(q, node_make(la, nla, cur.ch, cur.nch, i, 1));
            ++task_count;
        }
        if (nra && !has_all_zero_clause(ra, nra)) {                            /* NDP:915-925 */
            fq_push(q, node_make(ra, nra, cur.ch, cur.nch, -i, 1));
            ++task_count;
        }
        node_free(&cur);
        ++iterations;                                                          /* NDP:926 */
        if (max_queues <= 0 && iterations >= max_iterations) break;            /* NDP:935-936 */
    }
    free(la);
    free(ra);
    if (initial_queue_size && *initial_queue_size == 0) *initial_queue_size = q->tail - q->head; /* NDP:939-941 */
    if (iterations_out) *iterations_out = iterations;
    if (task_count_out) *task_count_out = task_count;
    if (clause_evals) *clause_evals = evals;
    return q;
}

size_t orc_frontier_size(const orc_frontier* f) { return f->tail - f->head; }
size_t orc_frontier_nclauses(const orc_frontier* f, size_t k) { return f->v[f->head + k].n; }
size_t orc_frontier_nchoices(const orc_frontier* f, size_t k) { return f->v[f->head + k].nch; }
void orc_frontier_copy(const orc_frontier* f, size_t k, int32_t* clauses3, int32_t* choices) {
    const node_t* nd = &f->v[f->head + k];
    if (clauses3) memcpy(clauses3, nd->cl, nd->n * 3 * sizeof(int32_t));
    if (choices) memcpy(choices, nd->ch, nd->nch * sizeof(int32_t));
}
void orc_frontier_free(orc_frontier* f) {
    if (!f) return;
    for (size_t k = f->head; k < f->tail; ++k) node_free(&f->v[k]);
    free(f->v);
    free(f);
}

/* ------------------------------------------------------------------------------------------ */
/* Satisfy_iterative -- NDP-5_6_7.cpp:784-878.  LIFO over materialised clause sets.  LA is
 * pushed before RA (NDP:837-842 then 857-862) so RA (i := false) is explored first; an empty
 * LA is a model immediately (NDP:843-852), checked before RA (NDP:863-872). */
static uint64_t g_dfs_stats[8];
/* design statistics of the last orc_dfs call (not part of the reference's behaviour):
 * [0] nodes whose two children were both pushed, [1] nodes whose variable came from a
 * two-zero (unit) clause, [2] max stack height, [3] sum of |LA| over pushed LA siblings of [0],
 * [4] max |A| seen, [5] leaves (nodes that pushed nothing) */
const uint64_t* orc_last_dfs_stats(void) { return g_dfs_stats; }

int orc_dfs(const int32_t* c3, size_t n, int first_assignment, int32_t* model, size_t cap, size_t* model_len,
            uint64_t* nodes_out, uint64_t* clause_evals) {
    memset(g_dfs_stats, 0, sizeof g_dfs_stats);
    size_t scap = 64, sp = 0;
    node_t* st = (node_t*)malloc(scap * sizeof(node_t));
    st[sp++] = node_make(c3, n, NULL, 0, 0, 0);                                /* NDP:793-795 */
    int models = 0;
    uint64_t nodes = 0, evals = 0;
    size_t scratch_cap = n ? n : 1;
    int32_t* la = (int32_t*)malloc(scratch_cap * 3 * sizeof(int32_t));
    int32_t* ra = (int32_t*)malloc(scratch_cap * 3 * sizeof(int32_t));
    if (model_len) *model_len = 0;
#define RECORD_MODEL(CH, NCH, LIT, HAS)                                                     \
    do {                                                                                    \
        if (models == 0) {                                                                  \
            size_t len_ = (NCH) + ((HAS) ? 1 : 0);                                          \
            if (model_len) *model_len = len_;                                               \
            if (model) {                                                                    \
                for (size_t j_ = 0; j_ < (NCH) && j_ < cap; ++j_) model[j_] = (CH)[j_];     \
                if ((HAS) && (NCH) < cap) model[(NCH)] = (LIT);                             \
            }                                                                               \
        }                                                                                   \
        ++models;                                                                           \
    } while (0)
    while (sp) {                                                               /* NDP:801 */
        ++nodes;
        node_t cur = st[--sp];                                                 /* NDP:805-806 */
        int i = orc_choice(cur.cl, cur.n);                                     /* NDP:816 */
        if (i == 0) {                                                          /* NDP:817-827 */
            RECORD_MODEL(cur.ch, cur.nch, 0, 0);
            node_free(&cur);
            if (first_assignment) break;
            continue;
        }
        size_t nla, nra;
        int cla, cra;
        orc_resolution_step(cur.cl, cur.n, i, 1, la, &nla, ra, &nra, &cla, &cra); /* NDP:831 */
        evals += cur.n;
        {
            int unit = 0;
            for (size_t k = 0; k < cur.n && !unit; ++k) unit = abs_last_nonzero_if(cur.cl + 3 * k, 2) != 0;
            g_dfs_stats[1] += (uint64_t)unit;
            int open_la = nla && !cla, open_ra = nra && !cra;
            if (open_la && open_ra) { ++g_dfs_stats[0]; g_dfs_stats[3] += nla; }
            if (!open_la && !open_ra) ++g_dfs_stats[5];
            if (cur.n > g_dfs_stats[4]) g_dfs_stats[4] = cur.n;
        }
        int stop = 0;
        if (sp + 2 > scap) { scap *= 2; st = (node_t*)realloc(st, scap * sizeof(node_t)); }
        if (nla && !cla) {                                                     /* NDP:837-842 */
            st[sp++] = node_make(la, nla, cur.ch, cur.nch, i, 1);
        } else if (!nla) {                                                     /* NDP:843-852 */
            RECORD_MODEL(cur.ch, cur.nch, i, 1);
            if (first_assignment) stop = 1;
        }
        if (!stop) {
            if (nra && !cra) {                                                 /* NDP:857-862 */
                st[sp++] = node_make(ra, nra, cur.ch, cur.nch, -i, 1);
            } else if (!nra) {                                                 /* NDP:863-872 */
                RECORD_MODEL(cur.ch, cur.nch, -i, 1);
                if (first_assignment) stop = 1;
            }
        }
        node_free(&cur);
        if (sp > g_dfs_stats[2]) g_dfs_stats[2] = sp;
        if (stop) break;
    }
#undef RECORD_MODEL
    while (sp) node_free(&st[--sp]);
    free(st);
    free(la);
    free(ra);
    if (nodes_out) *nodes_out = nodes;
    if (clause_evals) *clause_evals = evals;
    return models;
}
