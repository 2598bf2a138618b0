// C interface of the SEXP value model in tests/stubs/Rcpp.h, for the Python harness (tests/rstub.py): build the arguments
// of a .Call and read the returned object.  Test infrastructure only.
#include <Rcpp.h>
#include <cstring>

extern "C" {

SEXP rstub_real(const double* v, long n) { SEXP s = rstub_alloc(RSTUB_REAL); s->real.assign(v, v + n); return s; }
SEXP rstub_matrix(const double* v, int nrow, int ncol)           // column-major, as R stores it
{
    SEXP s = rstub_alloc(RSTUB_REAL);
    s->real.assign(v, v + (size_t)nrow * (size_t)ncol);
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP rstub_int(const int* v, long n) { SEXP s = rstub_alloc(RSTUB_INT); s->integer.assign(v, v + n); return s; }
SEXP rstub_lgl(int v) { SEXP s = rstub_alloc(RSTUB_LGL); s->integer.push_back(v != 0); return s; }
SEXP rstub_list(void) { return rstub_alloc(RSTUB_LIST); }
void rstub_list_set(SEXP l, const char* name, SEXP v) { l->set(name, v); }

int rstub_type(SEXP s) { return s ? s->type : -1; }
long rstub_length(SEXP s) { return s->type == RSTUB_REAL ? (long)s->real.size() : s->type == RSTUB_LIST || s->type == RSTUB_S4 ? (long)s->items.size() : (long)s->integer.size(); }
const double* rstub_real_ptr(SEXP s) { return s->real.data(); }
const int* rstub_int_ptr(SEXP s) { return s->integer.data(); }
const char* rstub_text(SEXP s) { return s->text.c_str(); }
SEXP rstub_get(SEXP s, const char* name) { return s->find(name); }
const char* rstub_name(SEXP s, long i) { return s->items[(size_t)i].first.c_str(); }

}
