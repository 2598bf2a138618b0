/* Recording fake of libb200admm.so for tests/test_rglue_cpu.py: every entry point stores the arguments it was called with
 * and returns a canned result, so that the marshalling of r-pkg/src/b200_glue.cpp (argument order, types, the dgCMatrix
 * slots, ownership) can be executed without a GPU.  Test infrastructure only -- never linked into the product. */
#include <b200admm.h>
#include <stdlib.h>
#include <string.h>

struct fake_record {
    char entry[32];
    long long n, p; int dtype; const void* x; const void* y;
    double lambda[8]; int nlambda_given, nlambda; double lmin_ratio; int standardize, intercept;
    double alpha; int nthread;
    int maxit; double eps_abs, eps_rel, rho;
    int frees_path, frees_dense;
    int fail;                    /* next solver call returns B200ADMM error with this code */
} g_rec;

struct fake_record* fake_record(void) { return &g_rec; }
void fake_reset(void) { memset(&g_rec, 0, sizeof g_rec); }

static void rec_common(const char* entry, const b200admm_data* d, const b200admm_opts* o)
{
    strncpy(g_rec.entry, entry, sizeof g_rec.entry - 1);
    g_rec.n = d->n; g_rec.p = d->p; g_rec.dtype = d->dtype; g_rec.x = d->x; g_rec.y = d->y;
    g_rec.maxit = o->maxit; g_rec.eps_abs = o->eps_abs; g_rec.eps_rel = o->eps_rel; g_rec.rho = o->rho;
}
static void rec_path(const double* lam, int ng, int nl, double lmr, int st, int ic)
{
    int k;
    g_rec.nlambda_given = ng; g_rec.nlambda = nl; g_rec.lmin_ratio = lmr; g_rec.standardize = st; g_rec.intercept = ic;
    for (k = 0; k < ng && k < 8; k++) g_rec.lambda[k] = lam[k];
}
/* canned path: nl columns, (p + 1) rows; column k holds the intercept k + 0.5 and coefficient j = k + 1 with value -(k + 1) */
static void canned_path(b200admm_path* out, int nl, long long nrow, int with_intercept)
{
    int k; long long nnz = 0;
    memset(out, 0, sizeof *out);
    out->nlambda = nl; out->nrow = nrow;
    out->lambda = malloc(sizeof(double) * nl); out->niter = malloc(sizeof(int) * nl);
    out->colptr = malloc(sizeof(long long) * (nl + 1));
    out->rowidx = malloc(sizeof(int) * 2 * nl); out->val = malloc(sizeof(double) * 2 * nl);
    for (k = 0; k < nl; k++) {
        out->lambda[k] = 1.0 / (k + 1); out->niter[k] = 10 + k; out->colptr[k] = nnz;
        if (with_intercept) { out->rowidx[nnz] = 0; out->val[nnz] = k + 0.5; nnz++; }
        out->rowidx[nnz] = k + 1; out->val[nnz] = -(double)(k + 1); nnz++;
    }
    out->colptr[nl] = nnz;
}
static char g_err[64] = "";
static int maybe_fail(void)
{
    if (!g_rec.fail) return 0;
    strcpy(g_err, "fake library: requested failure");
    return g_rec.fail;
}

int b200admm_lasso(const b200admm_data* d, const double* lam, int ng, int nl, double lmr, int st, int ic,
                   const b200admm_opts* o, b200admm_path* out)
{
    rec_common("lasso", d, o); rec_path(lam, ng, nl, lmr, st, ic);
    if (maybe_fail()) return g_rec.fail;
    canned_path(out, ng > 0 ? ng : nl, d->p + 1, 1);
    return 0;
}
int b200admm_enet(const b200admm_data* d, const double* lam, int ng, int nl, double lmr, int st, int ic, double alpha,
                  const b200admm_opts* o, b200admm_path* out)
{
    rec_common("enet", d, o); rec_path(lam, ng, nl, lmr, st, ic); g_rec.alpha = alpha;
    if (maybe_fail()) return g_rec.fail;
    canned_path(out, ng > 0 ? ng : nl, d->p + 1, 1);
    return 0;
}
int b200admm_parlasso(const b200admm_data* d, const double* lam, int ng, int nl, double lmr, int st, int ic, int nthread,
                      const b200admm_opts* o, b200admm_path* out)
{
    rec_common("parlasso", d, o); rec_path(lam, ng, nl, lmr, st, ic); g_rec.nthread = nthread;
    if (maybe_fail()) return g_rec.fail;
    canned_path(out, ng > 0 ? ng : nl, d->p + 1, 1);
    return 0;
}
int b200admm_bp(const b200admm_data* d, const b200admm_opts* o, b200admm_path* out)
{
    rec_common("bp", d, o);
    if (maybe_fail()) return g_rec.fail;
    canned_path(out, 1, d->p, 0);
    return 0;
}
int b200admm_lad(const b200admm_data* d, int ic, const b200admm_opts* o, b200admm_dense* out)
{
    long long j;
    rec_common("lad", d, o); g_rec.intercept = ic;
    if (maybe_fail()) return g_rec.fail;
    memset(out, 0, sizeof *out);
    out->len = d->p + 1; out->beta = malloc(sizeof(double) * out->len); out->niter = 77;
    for (j = 0; j < out->len; j++) out->beta[j] = 0.25 * j;
    return 0;
}
void b200admm_free_path(b200admm_path* out)
{
    g_rec.frees_path++;
    free(out->lambda); free(out->niter); free(out->colptr); free(out->rowidx); free(out->val);
    memset(out, 0, sizeof *out);
}
void b200admm_free_dense(b200admm_dense* out) { g_rec.frees_dense++; free(out->beta); memset(out, 0, sizeof *out); }
const char* b200admm_last_error(void) { return g_err; }
