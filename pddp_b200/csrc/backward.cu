// Backward Riccati pass (ref: pddp/controllers/ilqr.py:489-674, default V_zz_reg=False branch;
// pddp/utils/constraint.py:150-266 for the box-constrained feed-forward term, nu == 1).
//
// Two kernels, both sequential in time from t = N-1 down to 0 with V_z / V_zz resident on chip:
//   * backward_thread_kernel<NZ>: one THREAD per problem, everything in registers; with the
//     PDDP_BATCH_INNER layout every load/store is coalesced across problems.  Used for the
//     small-state configs (pendulum nz=2, cartpole nz=4: the "backward-pass bandwidth study").
//   * backward_warp_kernel: one WARP per problem, V_zz / F_z / V_zz.F_z in shared memory, lanes
//     tile the nz x nz outputs; a problem's per-step record is contiguous (PDDP_PROBLEM_MAJOR) so
//     the warp streams it with coalesced loads.  Used for nz = 14 (UT-Cholesky cartpole) and
//     nz = 42 (full-covariance double cartpole).
// Roofline: HBM -- per trajectory-step 2nz^2+2nz*nu+nz+nu+nu^2 elements read, nu+nu*nz written.
#include "core.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace pddp {

// Scalar box-QP (ref: pddp/utils/constraint.py:150-266 for a 1 x 1 problem), in closed form.
//
// What the reference's projected-Newton loop does when n = 1 (Q > 0), step by step:
//   it 0: x = clamp(x0); g = Q x + c.  x on a bound with the gradient pointing outward -> result 6 (all clamped),
//         x is returned.  |g| < 1e-8 -> result 5, x is returned.  Otherwise the Newton target is
//         xs = -(c / chol) / chol and the line search tries xc = clamp(x + step (xs - x)), step = 1, 0.6, 0.36, ...
//         With theta = (xc - x) / (xs - x) in (0, 1] the Armijo ratio of a clamped candidate is
//         theta (1 - theta / 2) / step: it passes at step 1 unless the bound cuts the step below ~0.106 of its
//         length, and then it passes at the first step <= ~10 theta, where the candidate is STILL the bound.  An
//         unclamped candidate (step < theta) has ratio 1 - step / 2 >= 0.5.  So iteration 0 always ends on
//         x1 = clamp(x + (xs - x)).
//   it 1: improvement below tol |f| -> result 4 with the free set of iteration 0 (free).  Otherwise the point is
//         re-classified: on a bound with outward gradient -> result 6, not free; else interior, where g is rounding
//         noise: result 5, or one more noise-sized Newton step that ends in result 2 or 4 -- x moves by a few ulp.
// Every one of these exits is a success (result >= 1) and returns x1 up to rounding, so the loop collapses to the
// expressions below (same operation order as the reference for x1).  Failures: Q <= 0 or NaN -> -1 (potrf
// raises); non-finite c -> the reference spins 100 iterations on NaNs and returns 0.  The loop form (kept in
// backward_nu.cu for n > 1) was 75 % of this kernel's executed instructions: lanes of a warp run different
// iteration counts and the warp pays the union (profiles/r2_summary.md).
template <class T>
__device__ __forceinline__ int boxqp1(T x0, T Q, T c, T lo, T hi, T& x, bool& is_free) {
    const T min_grad = T(1e-8), tol = T(1e-8);
    x = clampv(x0, lo, hi);
    if (isinf(x)) x = T(0);
    is_free = true;
    if (!isfinite(c) || !isfinite(x)) return 0;
    const T g = Q * x + c;
    if ((x == lo && g > T(0)) || (x == hi && g < T(0))) { is_free = false; return 6; }
    if (!(Q > T(0))) return -1;
    if (fabs(g) < min_grad) return 5;
    const T f0 = T(0.5) * x * Q * x + x * c;
    const T chol = jsqrt(Q);
    const T search = -((c / chol) / chol) - x;
    x = clampv(x + search, lo, hi);
    const T f1 = T(0.5) * x * Q * x + x * c;
    if ((f0 - f1) < tol * fabs(f0)) return 4;
    const T g1 = Q * x + c;
    if ((x == lo && g1 > T(0)) || (x == hi && g1 < T(0))) { is_free = false; return 6; }
    return 4;
}

// gains for nu == 1.  Returns false where the reference raises (-> NOT_PD).
template <class T>
__device__ __forceinline__ bool gains1(T Quu, T Qu, T reg, bool bounded, T lo, T hi, T warm, T& k,
                                       T& inv /* K = -Q_uz * inv */) {
    if (!isfinite(Quu)) return false;                    // linalg.eig raises on NaN/Inf
    T e = Quu < T(0) ? T(1e-12) : Quu;                    // ref: ilqr.py:633
    e += reg;
    if (!bounded) {
        inv = T(1) / e;
        k = -inv * Qu;
        return !(k != k);
    }
    bool is_free;
    int result = boxqp1(warm, e, Qu, lo, hi, k, is_free);
    if (result < 1) return false;
    T ch = jsqrt(e);
    inv = is_free ? (T(1) / ch) / ch : T(0);
    return true;
}

// ------------------------------------------------------------------------------------------
// One time step's inputs of one problem, held in registers.  The time loop is a dependent chain
// (V_z, V_zz), but its INPUTS are not: the record of step t-DEPTH is requested before step t is
// computed, so every thread always has DEPTH records (17 floats at nz = 2) in flight -- that, times
// the resident threads, is what keeps HBM busy in a thread-per-problem scan.
template <class T, int NZ>
struct BackwardRec {
    T Fz[NZ][NZ], Fu[NZ], Lz[NZ], Luz[NZ], Lzz[NZ][NZ], Lu, Luu, U;
};
template <class T, int NZ>
__device__ __forceinline__ void backward_load(const BackwardArgs<T>& a, int b, int t, bool bounded, BackwardRec<T, NZ>& r) {
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
        r.Fu[i] = __ldg(a.F_u + a.lFu.at(b, t, i));
        r.Lz[i] = __ldg(a.L_z + a.lLz.at(b, t, i));
        r.Luz[i] = __ldg(a.L_uz + a.lLuz.at(b, t, i));
#pragma unroll
        for (int j = 0; j < NZ; ++j) {
            r.Fz[i][j] = __ldg(a.F_z + a.lFz.at(b, t, i * NZ + j));
            r.Lzz[i][j] = __ldg(a.L_zz + a.lLzz.at(b, t, i * NZ + j));
        }
    }
    r.Lu = __ldg(a.L_u + a.lLu.at(b, t, 0));
    r.Luu = __ldg(a.L_uu + a.lLuu.at(b, t, 0));
    r.U = bounded ? __ldg(a.U + a.lU.at(b, t, 0)) : T(0);
}

template <class T, int NZ>
__global__ void __launch_bounds__(128, (sizeof(T) == 4 && NZ <= 2) ? 5 : 2) backward_thread_kernel(const BackwardArgs<T> a) {
    constexpr bool DEEP = NZ <= 2;          // two records in flight where they fit in registers
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    if (a.active && a.active[b] == 0) return;
    const bool bounded = a.u_min != nullptr && a.u_max != nullptr;
    const T reg = (T)a.mu[b];
    T lo = T(0), hi = T(0);
    if (bounded) { lo = a.u_min[0]; hi = a.u_max[0]; }
    BackwardRec<T, NZ> c, nx1, nx2;
    backward_load<T, NZ>(a, b, a.N - 1, bounded, c);
    if (DEEP && a.N >= 2) backward_load<T, NZ>(a, b, a.N - 2, bounded, nx1);
    T v[NZ], V[NZ][NZ];
#pragma unroll
    for (int i = 0; i < NZ; ++i) {
        v[i] = a.L_z[a.lLz.at(b, a.N, i)];
#pragma unroll
        for (int j = 0; j < NZ; ++j) V[i][j] = a.L_zz[a.lLzz.at(b, a.N, i * NZ + j)];
    }
    T k_next = T(0);
    bool ok = true;
#pragma unroll 1
    for (int t = a.N - 1; t >= 0; --t) {
        {
            if (DEEP) { if (t >= 2) backward_load<T, NZ>(a, b, t - 2, bounded, nx2); }
            else if (t >= 1) backward_load<T, NZ>(a, b, t - 1, bounded, nx1);
            T W[NZ][NZ], wu[NZ];                               // W = V Fz, wu = V Fu
#pragma unroll
            for (int i = 0; i < NZ; ++i) {
                T su = T(0);
#pragma unroll
                for (int kk = 0; kk < NZ; ++kk) su += V[i][kk] * c.Fu[kk];
                wu[i] = su;
#pragma unroll
                for (int j = 0; j < NZ; ++j) {
                    T sacc = T(0);
#pragma unroll
                    for (int kk = 0; kk < NZ; ++kk) sacc += V[i][kk] * c.Fz[kk][j];
                    W[i][j] = sacc;
                }
            }
            T Qz[NZ], Quz[NZ], Qzz[NZ][NZ];
            T Qu = c.Lu, Quu = c.Luu;
#pragma unroll
            for (int kk = 0; kk < NZ; ++kk) {
                Qu += c.Fu[kk] * v[kk];
                Quu += c.Fu[kk] * wu[kk];
            }
#pragma unroll
            for (int i = 0; i < NZ; ++i) {
                T sz = c.Lz[i], suz = c.Luz[i];
#pragma unroll
                for (int kk = 0; kk < NZ; ++kk) {
                    sz += c.Fz[kk][i] * v[kk];
                    suz += c.Fu[kk] * W[kk][i];
                }
                Qz[i] = sz;
                Quz[i] = suz;
#pragma unroll
                for (int j = 0; j < NZ; ++j) {
                    T sacc = c.Lzz[i][j];
#pragma unroll
                    for (int kk = 0; kk < NZ; ++kk) sacc += c.Fz[kk][i] * W[kk][j];
                    Qzz[i][j] = sacc;
                }
            }
            T kt, inv;
            ok = gains1(Quu, Qu, reg, bounded, lo - c.U, hi - c.U, k_next, kt, inv);
            T Kt[NZ];
#pragma unroll
            for (int i = 0; i < NZ; ++i) {
                Kt[i] = -Quz[i] * inv;
                if (Kt[i] != Kt[i]) ok = false;
            }
            if (!ok) break;
            k_next = kt;
            a.k[a.lk.at(b, t, 0)] = kt;
#pragma unroll
            for (int i = 0; i < NZ; ++i) a.K[a.lK.at(b, t, i)] = Kt[i];
            // value update with the UN-regularised Q_uu (ref: ilqr.py:664-672)
#pragma unroll
            for (int i = 0; i < NZ; ++i) v[i] = Qz[i] + Kt[i] * Qu + Kt[i] * Quu * kt + Quz[i] * kt;
#pragma unroll
            for (int i = 0; i < NZ; ++i)
#pragma unroll
                for (int j = 0; j < NZ; ++j)
                    V[i][j] = T(0.5) * (Qzz[i][j] + Qzz[j][i]) + Kt[i] * Quu * Kt[j] + Kt[i] * Quz[j] + Quz[i] * Kt[j];
        }
        c = nx1;
        if (DEEP) nx1 = nx2;
    }
    a.status[b] = ok ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
// A TEAM of threads per problem; dynamic smem: per team 3*nz*nz + 6*nz elements.
//   TEAM = 32 : a warp per problem, 4 problems per CTA (nz = 14: 196 outputs per product).
//   TEAM = 256: a CTA per problem (nz = 42: 1764 outputs x 42 MACs per product; with one warp the
//               time loop is a chain of LDS->FMA latencies, 8 warps split every product 8 ways).
template <int TEAM>
__device__ __forceinline__ void team_sync() {
    if (TEAM == 32) __syncwarp(); else __syncthreads();
}
template <int TEAM>
__device__ __forceinline__ bool team_any(bool x) {
    if (TEAM == 32) return __any_sync(0xffffffffu, x);
    return __syncthreads_or(x) != 0;
}
// elements of shared memory per team: three nz x LD matrices + six LD vectors
__host__ __device__ inline int backward_team_elems(int nz) {
    const int LD = (nz + 3) & ~3;
    return 3 * nz * LD + 6 * LD;
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
    const double2 t0 = *reinterpret_cast<const double2*>(p), t1 = *reinterpret_cast<const double2*>(p + 2);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
}
__device__ __forceinline__ void load2(const float* p, float (&v)[2]) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
}
__device__ __forceinline__ void load2(const double* p, double (&v)[2]) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x; v[1] = t.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
// NZC > 0: encoded state size known at compile time (nz = 14: UT-Cholesky cartpole, nz = 42: full-covariance double
// cartpole) -- the strip loops unroll and the index divisions fold away.
template <class T, int TEAM, int NZC = 0>
// (min blocks per SM 7: 4 096 problems at 4 per CTA, or 1 024 at one CTA each, are 6.9 CTAs per SM -- with the
// 96 / 40 registers the compiler took by itself 5 / 6 were resident and a second, nearly empty wave cost 15 - 38 %)
__global__ void __launch_bounds__(TEAM == 32 ? 128 : TEAM, sizeof(T) == 4 ? 7 : 1) backward_warp_kernel(const BackwardArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x % TEAM, warp = threadIdx.x / TEAM, wpb = blockDim.x / TEAM;
    const int b = blockIdx.x * wpb + warp;
    if (b >= a.B) return;
    if (a.active && a.active[b] == 0) return;
    // rows padded to LD (a multiple of 4) so that a thread can own a 1 x 4 strip of an output row and
    // read its B operand with one 16-byte LDS per 4 FMAs (an element per thread needs 2 LDS per FMA,
    // and at nz = 42 the pass was bound by shared-memory loads)
    // Register tile: 2 rows x 4 columns per thread.  Both products read their A operand DOWN a column of a row-major
    // matrix (W = V Fz through VT[kk][i] = V[i][kk]: V is kept TRANSPOSED -- it is symmetric after the first step and
    // the terminal L_zz is stored transposed below --; Q_zz = Fz^T W through Fz[kk][i]), so the two A values of a tile
    // are one 8-byte LDS: 6 shared-memory wavefronts and 10 instructions per 8 FMAs instead of 5 and 6 per 4 (the
    // nz = 42 pass was bound by the shared-memory pipe, the nz = 14 pass needed two rounds of 1 x 4 strips per product).
    const int nz = NZC > 0 ? NZC : a.nz, nn = nz * nz, LD = (nz + 3) & ~3, NS4 = LD / 4, nl = nz * LD, NR2 = (nz + 1) / 2;
    const int per_team = backward_team_elems(nz);
    T* base = reinterpret_cast<T*>(smem_raw) + (size_t)warp * per_team;
    T *V = base, *Fz = base + nl, *W = base + 2 * nl;
    T *v = base + 3 * nl, *Fu = v + LD, *wu = Fu + LD, *Quz = wu + LD, *Qz = Quz + LD, *Kt = Qz + LD;
    for (int e = lane; e < 3 * nl; e += TEAM) base[e] = T(0);      // the padding columns stay zero
    team_sync<TEAM>();
    const bool bounded = a.u_min != nullptr && a.u_max != nullptr;
    const T reg = (T)a.mu[b];
    T lo = T(0), hi = T(0);
    if (bounded) { lo = a.u_min[0]; hi = a.u_max[0]; }

    for (int e = lane; e < nn; e += TEAM) V[(e % nz) * LD + e / nz] = a.L_zz[a.lLzz.at(b, a.N, e)];   // V^T
    for (int e = lane; e < nz; e += TEAM) v[e] = a.L_z[a.lLz.at(b, a.N, e)];
    team_sync<TEAM>();
    T k_next = T(0);
    bool ok = true;
    for (int t = a.N - 1; t >= 0; --t) {
        for (int e = lane; e < nn; e += TEAM) Fz[(e / nz) * LD + e % nz] = a.F_z[a.lFz.at(b, t, e)];
        for (int e = lane; e < nz; e += TEAM) Fu[e] = a.F_u[a.lFu.at(b, t, e)];
        team_sync<TEAM>();
        // W = V Fz ; wu = V Fu
        for (int s4 = lane; s4 < NR2 * NS4; s4 += TEAM) {
            const int i0 = (s4 / NS4) * 2, j0 = (s4 % NS4) * 4;
            T acc[2][4] = {{T(0), T(0), T(0), T(0)}, {T(0), T(0), T(0), T(0)}};
            for (int kk = 0; kk < nz; ++kk) {
                T av[2], bv[4];
                load2(V + kk * LD + i0, av);                     // V[i0][kk], V[i0 + 1][kk]  (row i0 + 1 == nz: zero padding)
                load4(Fz + kk * LD + j0, bv);
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] += av[r] * bv[c];
            }
            store4(W + i0 * LD + j0, acc[0]);
            if (i0 + 1 < nz) store4(W + (i0 + 1) * LD + j0, acc[1]);
        }
        for (int i = lane; i < nz; i += TEAM) {
            T s = T(0);
            for (int kk = 0; kk < nz; ++kk) s += V[kk * LD + i] * Fu[kk];
            wu[i] = s;
        }
        team_sync<TEAM>();
        // Q_z, Q_uz (vectors), Q_u, Q_uu (scalars, every lane computes them redundantly)
        for (int i = lane; i < nz; i += TEAM) {
            T sz = a.L_z[a.lLz.at(b, t, i)], suz = a.L_uz[a.lLuz.at(b, t, i)];
            for (int kk = 0; kk < nz; ++kk) {
                sz += Fz[kk * LD + i] * v[kk];
                suz += Fu[kk] * W[kk * LD + i];
            }
            Qz[i] = sz;
            Quz[i] = suz;
        }
        T Qu = a.L_u[a.lLu.at(b, t, 0)], Quu = a.L_uu[a.lLuu.at(b, t, 0)];
        for (int kk = 0; kk < nz; ++kk) {
            Qu += Fu[kk] * v[kk];
            Quu += Fu[kk] * wu[kk];
        }
        team_sync<TEAM>();
        // Q_zz = L_zz + Fz^T W  -> overwrites V (V is dead once W and wu exist)
        for (int s4 = lane; s4 < NR2 * NS4; s4 += TEAM) {
            const int i0 = (s4 / NS4) * 2, j0 = (s4 % NS4) * 4;
            T acc[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    acc[r][c] = (i0 + r < nz && j0 + c < nz) ? a.L_zz[a.lLzz.at(b, t, (i0 + r) * nz + j0 + c)] : T(0);
            for (int kk = 0; kk < nz; ++kk) {
                T av[2], bv[4];
                load2(Fz + kk * LD + i0, av);
                load4(W + kk * LD + j0, bv);
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] += av[r] * bv[c];
            }
            store4(V + i0 * LD + j0, acc[0]);                    // Q_zz, row-major; symmetrised below
            if (i0 + 1 < nz) store4(V + (i0 + 1) * LD + j0, acc[1]);
        }
        T kt, inv;
        T ut = bounded ? a.U[a.lU.at(b, t, 0)] : T(0);
        ok = gains1(Quu, Qu, reg, bounded, lo - ut, hi - ut, k_next, kt, inv);
        bool bad = false;
        for (int i = lane; i < nz; i += TEAM) {
            T Ki = -Quz[i] * inv;
            Kt[i] = Ki;
            bad |= (Ki != Ki);
        }
        ok = ok && !team_any<TEAM>(bad);
        if (!ok) break;
        team_sync<TEAM>();
        k_next = kt;
        if (lane == 0) a.k[a.lk.at(b, t, 0)] = kt;
        for (int i = lane; i < nz; i += TEAM) {
            a.K[a.lK.at(b, t, i)] = Kt[i];
            v[i] = Qz[i] + Kt[i] * Qu + Kt[i] * Quu * kt + Quz[i] * kt;
        }
        // symmetrise Q_zz and add the gain terms, one lane owns both (i,j) and (j,i)
        for (int e = lane; e < nn; e += TEAM) {
            const int i = e / nz, j = e - i * nz;
            if (j < i) continue;
            T q = T(0.5) * (V[i * LD + j] + V[j * LD + i]);
            T val = q + Kt[i] * Quu * Kt[j] + Kt[i] * Quz[j] + Quz[i] * Kt[j];
            V[i * LD + j] = val;
            V[j * LD + i] = val;
        }
        team_sync<TEAM>();
    }
    if (lane == 0) a.status[b] = ok ? 0 : 1;
}

template <class T>
cudaError_t backward_pass(const BackwardArgs<T>& a, int layout, cudaStream_t s) {
    // action_size > 1 (eigen-clipping by Jacobi rotations, n-dimensional box QP): backward_nu.cu.
    // PDDP_FORCE_BACKWARD_NU=1 sends nu == 1 there too (tests compare the two kernels).
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("PDDP_FORCE_BACKWARD_NU"); forced = (e && e[0] == '1') ? 1 : 0; }
    if (a.nu > 1 || forced) return backward_pass_nu<T>(a, s);
    const int threads = 128;
    if (layout == LAYOUT_BATCH_INNER && (a.nz == 2 || a.nz == 4)) {
        const int grid = (a.B + threads - 1) / threads;
        if (a.nz == 2) backward_thread_kernel<T, 2><<<grid, threads, 0, s>>>(a);
        else backward_thread_kernel<T, 4><<<grid, threads, 0, s>>>(a);
        return cudaGetLastError();
    }
    const size_t per_team = (size_t)backward_team_elems(a.nz) * sizeof(T);
    if (a.nz >= 24) {                                   // a CTA per problem
        if (per_team > 227 * 1024) return cudaErrorInvalidValue;
        auto kern = a.nz == 42 ? backward_warp_kernel<T, 256, 42> : backward_warp_kernel<T, 256, 0>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_team);
        if (e != cudaSuccess) return e;
        kern<<<a.B, 256, per_team, s>>>(a);
        return cudaGetLastError();
    }
    int wpb = 4;
    while (wpb > 1 && per_team * wpb > 200 * 1024) wpb >>= 1;
    const size_t smem = per_team * wpb;
    auto kern = a.nz == 14 ? backward_warp_kernel<T, 32, 14> : backward_warp_kernel<T, 32, 0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(a.B + wpb - 1) / wpb, wpb * 32, smem, s>>>(a);
    return cudaGetLastError();
}
template cudaError_t backward_pass<float>(const BackwardArgs<float>&, int, cudaStream_t);
template cudaError_t backward_pass<double>(const BackwardArgs<double>&, int, cudaStream_t);

}  // namespace pddp
