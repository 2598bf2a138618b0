// Stand-in for <Rcpp.h> (R and Rcpp are not in this image): the part of the API surface that r-pkg/src/b200_glue.cpp uses,
// over a small value model of R's SEXP, so that the glue can be compiled AND executed by the tests
// (tests/test_rglue_cpu.py against a recording fake of the library, tests/test_gpu_rglue.py against the CUDA library).
// Semantics kept from R / Rcpp: vectors are handles on a shared SEXP (no copy when wrapping an argument), a matrix is a
// real vector with a dim attribute in column-major order, as<T>() converts length-one vectors, a list is addressed by
// name, an S4 object has named slots, stop() raises an R error that END_RCPP turns into a condition object.
#pragma once
#include <algorithm>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

typedef std::ptrdiff_t R_xlen_t;

enum { RSTUB_NIL = 0, RSTUB_LGL = 10, RSTUB_INT = 13, RSTUB_REAL = 14, RSTUB_LIST = 19, RSTUB_S4 = 25, RSTUB_ERROR = 99 };

struct SEXPREC {
    int type = RSTUB_NIL;
    std::vector<double> real;
    std::vector<int> integer;                                   // INT and LGL payload
    int nrow = 0, ncol = 0;                                     // dim attribute (matrices)
    std::vector<std::pair<std::string, SEXPREC*> > items;      // list elements / S4 slots
    std::string text;                                           // S4 class name / error message
    SEXPREC* find(const char* name) const
    {
        for (size_t i = 0; i < items.size(); i++) if (items[i].first == name) return items[i].second;
        return 0;
    }
    void set(const char* name, SEXPREC* v)
    {
        for (size_t i = 0; i < items.size(); i++) if (items[i].first == name) { items[i].second = v; return; }
        items.push_back(std::make_pair(std::string(name), v));
    }
};
typedef SEXPREC* SEXP;

inline SEXP rstub_alloc(int type) { SEXP s = new SEXPREC; s->type = type; return s; }   // never freed: test processes only

#define RcppExport extern "C"
#define BEGIN_RCPP try {
#define END_RCPP } catch (std::exception& e__) { SEXP c__ = rstub_alloc(RSTUB_ERROR); c__->text = e__.what(); return c__; } \
                 return rstub_alloc(RSTUB_NIL);

namespace Rcpp {

struct not_compatible : std::runtime_error { explicit not_compatible(const std::string& m) : std::runtime_error(m) {} };
inline void stop(const char* msg) { throw std::runtime_error(msg); }

template <class T> struct sexp_type;
template <> struct sexp_type<double> { static const int value = RSTUB_REAL; };
template <> struct sexp_type<int> { static const int value = RSTUB_INT; };

inline double scalar_of(SEXP s)
{
    if (!s) throw not_compatible("Expecting a single value: [extent=0]");
    if (s->type == RSTUB_REAL && s->real.size() == 1) return s->real[0];
    if ((s->type == RSTUB_INT || s->type == RSTUB_LGL) && s->integer.size() == 1) return (double)s->integer[0];
    throw not_compatible("Expecting a single value");
}
template <class T> T as(SEXP s) { return (T)scalar_of(s); }
template <> inline bool as<bool>(SEXP s) { return scalar_of(s) != 0.0; }

template <class T> struct named_value { std::string name; T value; };
struct Named {
    std::string name;
    explicit Named(const char* n) : name(n) {}
    template <class T> named_value<T> operator=(const T& v) const { named_value<T> r = {name, v}; return r; }
};

template <class T> class Vector {
    SEXP s_;
    std::vector<T>& data() const;
public:
    Vector() : s_(rstub_alloc(sexp_type<T>::value)) {}
    explicit Vector(SEXP s) : s_(s)                             // wraps, no copy; numeric(0) / integer coercion as in Rcpp
    {
        if (!s) throw not_compatible("NULL where a vector is expected");
        if (s->type == sexp_type<T>::value) return;
        if (sexp_type<T>::value == (int)RSTUB_REAL && s->type == (int)RSTUB_INT) {         // as.numeric() on an integer vector copies
            s_ = rstub_alloc(RSTUB_REAL);
            s_->real.assign(s->integer.begin(), s->integer.end());
            return;
        }
        throw not_compatible("not compatible with requested type");
    }
    explicit Vector(R_xlen_t n) : s_(rstub_alloc(sexp_type<T>::value)) { data().resize((size_t)n); }
    explicit Vector(int n) : s_(rstub_alloc(sexp_type<T>::value)) { data().resize((size_t)n); }
    template <class It> Vector(It a, It b) : s_(rstub_alloc(sexp_type<T>::value)) { data().assign(a, b); }
    T* begin() { return data().data(); }
    R_xlen_t size() const { return (R_xlen_t)data().size(); }
    T& operator[](R_xlen_t i) { return data()[(size_t)i]; }
    static Vector create(T a, T b) { Vector r(2); r[0] = a; r[1] = b; return r; }
    operator SEXP() const { return s_; }
};
template <> inline std::vector<double>& Vector<double>::data() const { return s_->real; }
template <> inline std::vector<int>& Vector<int>::data() const { return s_->integer; }
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;

class NumericMatrix {
    SEXP s_;
public:
    explicit NumericMatrix(SEXP s) : s_(s)
    {
        if (!s || s->type != RSTUB_REAL) throw not_compatible("not a numeric matrix");
        if ((size_t)s->nrow * (size_t)s->ncol != s->real.size()) throw not_compatible("not a matrix");
    }
    int nrow() const { return s_->nrow; }
    int ncol() const { return s_->ncol; }
    double* begin() { return s_->real.data(); }
};

inline SEXP wrap(SEXP s) { return s; }
inline SEXP wrap(int v) { SEXP s = rstub_alloc(RSTUB_INT); s->integer.push_back(v); return s; }
inline SEXP wrap(double v) { SEXP s = rstub_alloc(RSTUB_REAL); s->real.push_back(v); return s; }
template <class T> SEXP wrap(const Vector<T>& v) { return (SEXP)v; }

class S4 {
    SEXP s_;
public:
    struct Slot {
        SEXP obj; std::string name;
        template <class T> Slot& operator=(const T& v) { obj->set(name.c_str(), wrap(v)); return *this; }
    };
    explicit S4(const char* cls) : s_(rstub_alloc(RSTUB_S4)) { s_->text = cls; }
    Slot slot(const char* name) { Slot r = {s_, name}; return r; }
    operator SEXP() const { return s_; }
};
inline SEXP wrap(const S4& v) { return (SEXP)v; }

class List {
    SEXP s_;
    void add() {}
    template <class T, class... A> void add(const named_value<T>& nv, const A&... rest) { s_->set(nv.name.c_str(), wrap(nv.value)); add(rest...); }
public:
    struct Proxy { SEXP value; operator SEXP() const { return value; } };
    List() : s_(rstub_alloc(RSTUB_LIST)) {}
    explicit List(SEXP s) : s_(s) { if (!s || s->type != RSTUB_LIST) throw not_compatible("not a list"); }
    Proxy operator[](const char* name) const
    {
        SEXP v = s_->find(name);
        if (!v) throw std::runtime_error(std::string("Index out of bounds: [index='") + name + "'].");
        Proxy p = {v};
        return p;
    }
    template <class... A> static List create(const A&... a) { List l; l.add(a...); return l; }
    operator SEXP() const { return s_; }
};
template <class T> T as(const List::Proxy& p) { return as<T>(p.value); }

}  // namespace Rcpp
