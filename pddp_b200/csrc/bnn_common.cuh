// Shared pieces of the BNN path: network descriptor, small dense linear algebra on register arrays
// (Cholesky with the reference's escalating jitter, its differential, triangular solves).
#pragma once
#include "core.cuh"

namespace pddp {

template <class T>
struct BnnNet {
    int P, H0, H1;
    const T *W0T, *b0, *W1T, *b1, *W2T, *b2;   // transposed copies: W0T[K0][H0], W1T[H0][H1], W2T[H1][D]
    const T *mask0, *mask1, *eps0;             // [P,H0], [P,H1], [P,D]
    const T *mask0T, *mask1T;                  // transposed copies [H0,P], [H1,P] (coalesced when lanes = particles)
    const T *X_mean, *X_std_inv, *dX_mean, *dX_std;   // nullptr = 0 / 1 defaults (ref: modules.py:93-98)
    const T* eps_out;                                 // [P,D] output noise of the CURRENT step, or nullptr (use_predicted_std=False)
    int independent_noise;
};

// Upper Cholesky factor, U^T U = C + jitter*I, reading the upper triangle of C.
template <int D, class T>
__device__ __forceinline__ bool chol_upper(const T (&C)[D][D], T jitter, T (&U)[D][D]) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if (j < i) { U[i][j] = T(0); continue; }
            T s = C[i][j] + (i == j ? jitter : T(0));
#pragma unroll
            for (int k = 0; k < i; ++k) s -= U[k][i] * U[k][j];
            if (i == j) {
                if (!(s > T(0))) ok = false;
                U[i][i] = jsqrt(s);
            } else {
                U[i][j] = s / U[i][i];
            }
        }
    }
    return ok;
}

// ref: pddp/utils/encoding.py:536-564 -- jitter 1e-12 is always added, x10 until it exceeds 10.
template <int D, class T>
__device__ __forceinline__ bool chol_upper_jitter(const T (&C)[D][D], T (&U)[D][D]) {
    double jitter = 1e-12;
    while (true) {
        if (chol_upper<D, T>(C, (T)jitter, U)) return true;
        jitter *= 10.0;
        if (jitter > 10.0) return false;
    }
}

// Differential of the upper Cholesky factor: given U (C = U^T U) and a SYMMETRIC dC, returns
// dU = Phi(U^-T dC U^-1) U with Phi = upper triangle, halved diagonal.  This is the forward-mode
// counterpart of torch's (symmetrised) Cholesky backward.
template <int D, class T>
__device__ __forceinline__ void chol_upper_diff(const T (&U)[D][D], const T (&dC)[D][D], T (&dU)[D][D]) {
    T Y[D][D];   // Y = U^-T dC  (forward substitution, U^T is lower)
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            T s = dC[i][c];
#pragma unroll
            for (int k = 0; k < i; ++k) s -= U[k][i] * Y[k][c];
            Y[i][c] = s / U[i][i];
        }
    T A[D][D];   // A = Y U^-1  (row-wise: a U = y)
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            T s = Y[r][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= A[r][k] * U[k][j];
            A[r][j] = s / U[j][j];
        }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            if (c < r) { dU[r][c] = T(0); continue; }
            T s = T(0);
#pragma unroll
            for (int k = r; k <= c; ++k) s += (k == r ? T(0.5) * A[r][k] : A[r][k]) * U[k][c];
            dU[r][c] = s;
        }
}

// eps = delta U^-1   (row vector times inverse of upper-triangular U)   ref: modules.py:335-348
template <int D, class T>
__device__ __forceinline__ void solve_right_upper(const T (&U)[D][D], const T* delta, T* eps) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
        T s = delta[c];
#pragma unroll
        for (int r = 0; r < c; ++r) s -= eps[r] * U[r][c];
        eps[c] = s / U[c][c];
    }
}

// U = decode_covar_sqrt(z)   ref: pddp/utils/encoding.py:304-362
template <int D, int ENC, class T>
__device__ __forceinline__ bool load_factor(const T* z, T (&U)[D][D]) {
    if (ENC == ENC_FULL) {
        T C[D][D];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = 0; b < D; ++b) C[a][b] = T(0.5) * (z[D + a * D + b] + z[D + b * D + a]);
        return chol_upper_jitter<D, T>(C, U);
    }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) {
            if (ENC == ENC_UT) U[a][b] = b >= a ? z[D + tri<D>(a, b)] : T(0);
            else if (ENC == ENC_VAR) U[a][b] = a == b ? jsqrt(z[D + a]) : T(0);      // diag(var).sqrt()
            else if (ENC == ENC_STD) U[a][b] = a == b ? z[D + a] : T(0);             // diag(std)
            else U[a][b] = a == b ? T(1e-3) : T(0);
        }
    return true;
}

// z' = encode(M, C=Cov)  for FULL / UT / IGNORE    ref: pddp/utils/encoding.py:99-141
template <int D, int ENC, class T>
__device__ __forceinline__ bool encode_moments(const T* M, const T (&Cov)[D][D], T* zn, T (&Un)[D][D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) zn[i] = M[i];
    if (ENC == ENC_FULL) {
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = 0; b < D; ++b) zn[D + a * D + b] = Cov[a][b];
        return true;
    }
    if (ENC == ENC_UT) {
        bool ok = chol_upper_jitter<D, T>(Cov, Un);
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = a; b < D; ++b) zn[D + tri<D>(a, b)] = Un[a][b];
        return ok;
    }
    if (ENC == ENC_VAR || ENC == ENC_STD) {      // ref: encoding.py:131-134 (_V_from / _S_from of C)
#pragma unroll
        for (int a = 0; a < D; ++a) zn[D + a] = ENC == ENC_VAR ? Cov[a][a] : jsqrt(Cov[a][a]);
    }
    return true;
}

// ---- 1-D bulk copy (TMA unit, cp.async.bulk) global -> shared with an mbarrier, for kernels that stage one contiguous
// block per CTA: one request instead of a dependent chain of per-thread loads
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bytes: a multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr_u32(dst)), "l"(src), "r"(bytes), "r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "BULK_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 10000;\n\t"
        "@p bra.uni BULK_WAIT_DONE;\n\t"
        "bra.uni BULK_WAIT_LOOP;\n\t"
        "BULK_WAIT_DONE:\n\t}" ::"r"(smem_addr_u32(bar)), "r"(parity) : "memory");
}

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

}  // namespace pddp
