// Minimal stand-in for <Rcpp.h>: just enough of the API surface for r-pkg/src/b200_glue.cpp to be type-checked
// (tests/test_rglue_cpu.py).  Nothing here is functional R; bodies exist only so the translation unit links into an object.
#pragma once
#include <algorithm>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

typedef struct SEXPREC* SEXP;
typedef std::ptrdiff_t R_xlen_t;
#define RcppExport extern "C"
#define BEGIN_RCPP try {
#define END_RCPP } catch (std::exception&) { return (SEXP)0; } return (SEXP)0;

namespace Rcpp {

template <class T> T as(SEXP) { return T(); }
inline void stop(const char* msg) { throw std::runtime_error(msg); }

template <class T> struct named_value { std::string name; T value; };
struct Named {
    std::string name;
    explicit Named(const char* n) : name(n) {}
    template <class T> named_value<T> operator=(const T& v) const { return named_value<T>{name, v}; }
};

template <class T> class Vector {
    std::vector<T> v_;
public:
    Vector() {}
    explicit Vector(SEXP) {}
    explicit Vector(R_xlen_t n) : v_((size_t)n) {}
    explicit Vector(int n) : v_((size_t)n) {}
    template <class It> Vector(It a, It b) : v_(a, b) {}
    T* begin() { return v_.data(); }
    R_xlen_t size() const { return (R_xlen_t)v_.size(); }
    T& operator[](R_xlen_t i) { return v_[(size_t)i]; }
    static Vector create(T a, T b) { Vector r(2); r[0] = a; r[1] = b; return r; }
    operator SEXP() const { return (SEXP)0; }
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;

class NumericMatrix {
    std::vector<double> v_;
    int nr_ = 0, nc_ = 0;
public:
    explicit NumericMatrix(SEXP) {}
    int nrow() const { return nr_; }
    int ncol() const { return nc_; }
    double* begin() { return v_.data(); }
};

class List {
public:
    struct Proxy { operator SEXP() const { return (SEXP)0; } };
    List() {}
    explicit List(SEXP) {}
    Proxy operator[](const char*) const { return Proxy(); }
    template <class... A> static List create(const A&...) { return List(); }
    operator SEXP() const { return (SEXP)0; }
};
template <class T> T as(const List::Proxy&) { return T(); }

class S4 {
public:
    struct Slot { template <class T> Slot& operator=(const T&) { return *this; } };
    explicit S4(const char*) {}
    Slot slot(const char*) { return Slot(); }
    operator SEXP() const { return (SEXP)0; }
};

}  // namespace Rcpp
