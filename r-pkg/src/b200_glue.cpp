// src/b200_glue.cpp  -- replaces src/Lasso.cpp, Enet.cpp, ParLasso.cpp, LAD.cpp, BP.cpp
#include <Rcpp.h>
#include <b200admm.h>
using namespace Rcpp;

static b200admm_opts get_opts(List opts) {            // src/Lasso.cpp:59-63
    b200admm_opts o;
    o.maxit   = as<int>(opts["maxit"]);
    o.eps_abs = as<double>(opts["eps_abs"]);
    o.eps_rel = as<double>(opts["eps_rel"]);
    o.rho     = as<double>(opts["rho"]);
    return o;
}
static b200admm_data get_data(NumericMatrix& x, NumericVector& y) {
    b200admm_data d;
    d.n = x.nrow(); d.p = x.ncol();
    d.dtype = B200ADMM_F64_HOST;                       // R's REALSXP, column-major, read-only
    d.x = x.begin(); d.y = y.begin();
    return d;
}
static void check(int rc) { if (rc != 0) stop(b200admm_last_error()); }   // BEGIN_RCPP / END_RCPP semantics

// b200admm_path -> List(lambda, beta = dgCMatrix, niter)   (src/Lasso.cpp:131-135)
static List wrap_path(b200admm_path& P, bool with_lambda) {
    const int nl = P.nlambda;
    const R_xlen_t nnz = P.colptr[nl];
    S4 beta("dgCMatrix");
    IntegerVector pp(nl + 1), ii(nnz);
    NumericVector xx(nnz);
    for (int k = 0; k <= nl; k++) pp[k] = (int)P.colptr[k];
    std::copy(P.rowidx, P.rowidx + nnz, ii.begin());
    std::copy(P.val, P.val + nnz, xx.begin());
    beta.slot("p") = pp; beta.slot("i") = ii; beta.slot("x") = xx;
    beta.slot("Dim") = IntegerVector::create((int)P.nrow, nl);
    IntegerVector niter(P.niter, P.niter + nl);
    NumericVector lambda(P.lambda, P.lambda + nl);
    b200admm_free_path(&P);
    if (with_lambda) return List::create(Named("lambda") = lambda, Named("beta") = beta, Named("niter") = niter);
    return List::create(Named("beta") = beta, Named("niter") = niter[0]);
}

RcppExport SEXP admm_lasso(SEXP x_, SEXP y_, SEXP lambda_, SEXP nlambda_, SEXP lmin_ratio_,
                           SEXP standardize_, SEXP intercept_, SEXP opts_) {
BEGIN_RCPP
    NumericMatrix x(x_); NumericVector y(y_), lambda(lambda_);
    b200admm_data d = get_data(x, y);
    b200admm_opts o = get_opts(List(opts_));
    b200admm_path P;
    check(b200admm_lasso(&d, lambda.begin(), lambda.size(), as<int>(nlambda_), as<double>(lmin_ratio_),
                         as<bool>(standardize_), as<bool>(intercept_), &o, &P));
    return wrap_path(P, true);
END_RCPP
}

RcppExport SEXP admm_enet(SEXP x_, SEXP y_, SEXP lambda_, SEXP nlambda_, SEXP lmin_ratio_,
                          SEXP standardize_, SEXP intercept_, SEXP alpha_, SEXP opts_) {
BEGIN_RCPP
    NumericMatrix x(x_); NumericVector y(y_), lambda(lambda_);
    b200admm_data d = get_data(x, y);
    b200admm_opts o = get_opts(List(opts_));
    b200admm_path P;
    check(b200admm_enet(&d, lambda.begin(), lambda.size(), as<int>(nlambda_), as<double>(lmin_ratio_),
                        as<bool>(standardize_), as<bool>(intercept_), as<double>(alpha_), &o, &P));
    return wrap_path(P, true);
END_RCPP
}

RcppExport SEXP admm_parlasso(SEXP x_, SEXP y_, SEXP lambda_, SEXP nlambda_, SEXP lmin_ratio_,
                              SEXP standardize_, SEXP intercept_, SEXP nthread_, SEXP opts_) {
BEGIN_RCPP
    NumericMatrix x(x_); NumericVector y(y_), lambda(lambda_);
    b200admm_data d = get_data(x, y);
    b200admm_opts o = get_opts(List(opts_));
    b200admm_path P;
    check(b200admm_parlasso(&d, lambda.begin(), lambda.size(), as<int>(nlambda_), as<double>(lmin_ratio_),
                            as<bool>(standardize_), as<bool>(intercept_), as<int>(nthread_), &o, &P));
    return wrap_path(P, true);
END_RCPP
}

RcppExport SEXP admm_lad(SEXP x_, SEXP y_, SEXP intercept_, SEXP opts_) {
BEGIN_RCPP
    NumericMatrix x(x_); NumericVector y(y_);
    b200admm_data d = get_data(x, y);
    b200admm_opts o = get_opts(List(opts_));
    b200admm_dense D;
    check(b200admm_lad(&d, as<bool>(intercept_), &o, &D));
    NumericVector beta(D.beta, D.beta + D.len);
    const int niter = D.niter;
    b200admm_free_dense(&D);
    return List::create(Named("beta") = beta, Named("niter") = niter);      // src/LAD.cpp:44-45
END_RCPP
}

RcppExport SEXP admm_bp(SEXP x_, SEXP y_, SEXP opts_) {
BEGIN_RCPP
    NumericMatrix x(x_); NumericVector y(y_);
    b200admm_data d = get_data(x, y);
    b200admm_opts o = get_opts(List(opts_));
    b200admm_path P;
    check(b200admm_bp(&d, &o, &P));
    return wrap_path(P, false);                                             // src/BP.cpp:38-43
END_RCPP
}
