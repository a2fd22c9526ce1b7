// Forward-mode AD scalars used to fuse the linearisation into the dynamics / cost kernels.
//
// The reference obtains d z'/d[z,u] and the cost gradient/Hessian by reverse-mode autograd over
// replicated rows (pddp/utils/evaluation.py:134-288).  Here the same functions are evaluated once
// on "jets": Jet1<T,N> carries N first-order tangents, Jet2<T,N> additionally carries the packed
// upper triangle of the second-order tangents.  With N=2 and seeds (e_i, e_j) a Jet2 is a
// hyper-dual number whose h(0,1) component is exactly d2f/dz_i dz_j.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define PDDP_HD __host__ __device__ __forceinline__

namespace pddp {

// ---------------------------------------------------------------- plain scalar overloads
PDDP_HD float jsin(float x) { return sinf(x); }
PDDP_HD double jsin(double x) { return sin(x); }
PDDP_HD float jcos(float x) { return cosf(x); }
PDDP_HD double jcos(double x) { return cos(x); }
// sine and cosine of one argument with ONE range reduction
PDDP_HD void jsincos(float x, float& s, float& c) { sincosf(x, &s, &c); }
PDDP_HD void jsincos(double x, double& s, double& c) { sincos(x, &s, &c); }
PDDP_HD float jexp(float x) { return expf(x); }
PDDP_HD double jexp(double x) { return exp(x); }
PDDP_HD float jsqrt(float x) { return sqrtf(x); }
PDDP_HD double jsqrt(double x) { return sqrt(x); }
PDDP_HD float jvalue(float x) { return x; }
PDDP_HD double jvalue(double x) { return x; }

// ---------------------------------------------------------------- first order
template <class T, int N>
struct Jet1 {
    T v;
    T d[N];
    PDDP_HD Jet1() {}
    PDDP_HD Jet1(T value) : v(value) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = T(0);
    }
    PDDP_HD static Jet1 variable(T value, int dir) {
        Jet1 r(value);
#pragma unroll
        for (int i = 0; i < N; ++i) r.d[i] = (i == dir) ? T(1) : T(0);
        return r;
    }
};

template <class T, int N>
PDDP_HD Jet1<T, N> chain(const Jet1<T, N>& a, T f, T fp) {
    Jet1<T, N> r;
    r.v = f;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = fp * a.d[i];
    return r;
}
template <class T, int N>
PDDP_HD Jet1<T, N> operator+(const Jet1<T, N>& a, const Jet1<T, N>& b) {
    Jet1<T, N> r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <class T, int N>
PDDP_HD Jet1<T, N> operator-(const Jet1<T, N>& a, const Jet1<T, N>& b) {
    Jet1<T, N> r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <class T, int N>
PDDP_HD Jet1<T, N> operator-(const Jet1<T, N>& a) {
    Jet1<T, N> r;
    r.v = -a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
    return r;
}
template <class T, int N>
PDDP_HD Jet1<T, N> operator*(const Jet1<T, N>& a, const Jet1<T, N>& b) {
    Jet1<T, N> r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.v * b.d[i] + a.d[i] * b.v;
    return r;
}
template <class T, int N>
PDDP_HD Jet1<T, N> operator/(const Jet1<T, N>& a, const Jet1<T, N>& b) {
    Jet1<T, N> r;
    T inv = T(1) / b.v;
    r.v = a.v * inv;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <class T, int N> PDDP_HD Jet1<T, N> operator+(const Jet1<T, N>& a, T s) { Jet1<T, N> r = a; r.v += s; return r; }
template <class T, int N> PDDP_HD Jet1<T, N> operator+(T s, const Jet1<T, N>& a) { return a + s; }
template <class T, int N> PDDP_HD Jet1<T, N> operator-(const Jet1<T, N>& a, T s) { Jet1<T, N> r = a; r.v -= s; return r; }
template <class T, int N> PDDP_HD Jet1<T, N> operator-(T s, const Jet1<T, N>& a) { return (-a) + s; }
template <class T, int N> PDDP_HD Jet1<T, N> operator*(const Jet1<T, N>& a, T s) { return chain(a, a.v * s, s); }
template <class T, int N> PDDP_HD Jet1<T, N> operator*(T s, const Jet1<T, N>& a) { return a * s; }
template <class T, int N> PDDP_HD Jet1<T, N> operator/(const Jet1<T, N>& a, T s) { return a * (T(1) / s); }
template <class T, int N> PDDP_HD Jet1<T, N> operator/(T s, const Jet1<T, N>& a) { return Jet1<T, N>(s) / a; }
template <class T, int N> PDDP_HD Jet1<T, N> jsin(const Jet1<T, N>& a) { T s, c; jsincos(a.v, s, c); return chain(a, s, c); }
template <class T, int N> PDDP_HD Jet1<T, N> jcos(const Jet1<T, N>& a) { T s, c; jsincos(a.v, s, c); return chain(a, c, -s); }
template <class T, int N> PDDP_HD void jsincos(const Jet1<T, N>& a, Jet1<T, N>& sn, Jet1<T, N>& cs) {
    T s, c;
    jsincos(a.v, s, c);
    sn = chain(a, s, c);
    cs = chain(a, c, -s);
}
template <class T, int N> PDDP_HD Jet1<T, N> jexp(const Jet1<T, N>& a) { T e = jexp(a.v); return chain(a, e, e); }
template <class T, int N> PDDP_HD Jet1<T, N> jsqrt(const Jet1<T, N>& a) { T s = jsqrt(a.v); return chain(a, s, T(0.5) / s); }
template <class T, int N> PDDP_HD T jvalue(const Jet1<T, N>& a) { return a.v; }

// ---------------------------------------------------------------- second order
template <int N> PDDP_HD constexpr int tri(int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }

template <class T, int N>
struct Jet2 {
    static constexpr int NH = N * (N + 1) / 2;
    T v;
    T g[N];
    T h[NH];   // packed upper triangle, h[tri(i,j)] = d2/d_i d_j  (i <= j)
    PDDP_HD Jet2() {}
    PDDP_HD Jet2(T value) : v(value) {
#pragma unroll
        for (int i = 0; i < N; ++i) g[i] = T(0);
#pragma unroll
        for (int i = 0; i < NH; ++i) h[i] = T(0);
    }
};

// result of f(a) given f, f', f'' at a.v
template <class T, int N>
PDDP_HD Jet2<T, N> chain(const Jet2<T, N>& a, T f, T fp, T fpp) {
    Jet2<T, N> r;
    r.v = f;
#pragma unroll
    for (int i = 0; i < N; ++i) r.g[i] = fp * a.g[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j)
            r.h[tri<N>(i, j)] = fp * a.h[tri<N>(i, j)] + fpp * a.g[i] * a.g[j];
    return r;
}
template <class T, int N>
PDDP_HD Jet2<T, N> operator+(const Jet2<T, N>& a, const Jet2<T, N>& b) {
    Jet2<T, N> r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.g[i] = a.g[i] + b.g[i];
#pragma unroll
    for (int i = 0; i < Jet2<T, N>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
    return r;
}
template <class T, int N>
PDDP_HD Jet2<T, N> operator-(const Jet2<T, N>& a, const Jet2<T, N>& b) {
    Jet2<T, N> r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.g[i] = a.g[i] - b.g[i];
#pragma unroll
    for (int i = 0; i < Jet2<T, N>::NH; ++i) r.h[i] = a.h[i] - b.h[i];
    return r;
}
template <class T, int N>
PDDP_HD Jet2<T, N> operator-(const Jet2<T, N>& a) {
    Jet2<T, N> r;
    r.v = -a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.g[i] = -a.g[i];
#pragma unroll
    for (int i = 0; i < Jet2<T, N>::NH; ++i) r.h[i] = -a.h[i];
    return r;
}
template <class T, int N>
PDDP_HD Jet2<T, N> operator*(const Jet2<T, N>& a, const Jet2<T, N>& b) {
    Jet2<T, N> r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.g[i] = a.v * b.g[i] + a.g[i] * b.v;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j)
            r.h[tri<N>(i, j)] = a.v * b.h[tri<N>(i, j)] + a.h[tri<N>(i, j)] * b.v +
                                a.g[i] * b.g[j] + a.g[j] * b.g[i];
    return r;
}
template <class T, int N>
PDDP_HD Jet2<T, N> jrecip(const Jet2<T, N>& a) {
    T inv = T(1) / a.v;
    return chain(a, inv, -inv * inv, T(2) * inv * inv * inv);
}
template <class T, int N> PDDP_HD Jet2<T, N> operator/(const Jet2<T, N>& a, const Jet2<T, N>& b) { return a * jrecip(b); }
template <class T, int N> PDDP_HD Jet2<T, N> operator+(const Jet2<T, N>& a, T s) { Jet2<T, N> r = a; r.v += s; return r; }
template <class T, int N> PDDP_HD Jet2<T, N> operator+(T s, const Jet2<T, N>& a) { return a + s; }
template <class T, int N> PDDP_HD Jet2<T, N> operator-(const Jet2<T, N>& a, T s) { Jet2<T, N> r = a; r.v -= s; return r; }
template <class T, int N> PDDP_HD Jet2<T, N> operator-(T s, const Jet2<T, N>& a) { return (-a) + s; }
template <class T, int N> PDDP_HD Jet2<T, N> operator*(const Jet2<T, N>& a, T s) { return chain(a, a.v * s, s, T(0)); }
template <class T, int N> PDDP_HD Jet2<T, N> operator*(T s, const Jet2<T, N>& a) { return a * s; }
template <class T, int N> PDDP_HD Jet2<T, N> operator/(const Jet2<T, N>& a, T s) { return a * (T(1) / s); }
template <class T, int N> PDDP_HD Jet2<T, N> operator/(T s, const Jet2<T, N>& a) { return jrecip(a) * s; }
template <class T, int N> PDDP_HD Jet2<T, N> jsin(const Jet2<T, N>& a) { T s, c; jsincos(a.v, s, c); return chain(a, s, c, -s); }
template <class T, int N> PDDP_HD Jet2<T, N> jcos(const Jet2<T, N>& a) { T s, c; jsincos(a.v, s, c); return chain(a, c, -s, -c); }
template <class T, int N> PDDP_HD void jsincos(const Jet2<T, N>& a, Jet2<T, N>& sn, Jet2<T, N>& cs) {
    T s, c;
    jsincos(a.v, s, c);
    sn = chain(a, s, c, -s);
    cs = chain(a, c, -s, -c);
}
template <class T, int N> PDDP_HD Jet2<T, N> jexp(const Jet2<T, N>& a) { T e = jexp(a.v); return chain(a, e, e, e); }
template <class T, int N> PDDP_HD Jet2<T, N> jsqrt(const Jet2<T, N>& a) {
    T s = jsqrt(a.v);
    return chain(a, s, T(0.5) / s, T(-0.25) / (s * a.v));
}
template <class T, int N> PDDP_HD T jvalue(const Jet2<T, N>& a) { return a.v; }

}  // namespace pddp
