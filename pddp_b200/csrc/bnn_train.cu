// BNN training on the device: every step of ParticlesBNNDynamicsModel.fit
// (pddp/models/bnn/modules.py:131-198) -- forward with a fresh concrete / Bernoulli dropout mask per
// (row, unit) (modules.py:462-483, 550-583 with resample=True), the loss
//     -gaussian_log_likelihood(dx, mean, exp(log_std)).mean() + reg_scale * regularization() / N
// (losses.py:20-38; modules.py:434-447, 517-530, 749-766), its gradient, and the Adam(amsgrad) update --
// as two kernels per step, no host round trip, no atomics on the gradient path:
//
//   train_rows_kernel    CTA = ROWS mini-batch rows, thread = hidden unit.  Forward through the three layers
//                        (weights read coalesced from their transposed copies), the per-row loss and the
//                        whole per-row backward chain (dout -> dh1 -> dpre1 -> dh0 -> dpre0, and d/dlogit_p of
//                        the concrete masks).  Writes activations and deltas [batch, width].
//   train_update_kernel  thread = ONE parameter element: its gradient is a dot product over the batch rows
//                        (delta[:, j] . activation[:, k]), plus the regulariser's term; the thread applies the
//                        Adam(amsgrad) update to its element and refreshes the transposed copy.
//
// The network is small (6 -> 200 -> 200 -> 8, batch 128: 33 MFLOP per step) and the 2000-step loop is
// strictly sequential, so the design goal is latency per step (two launches, everything L2 resident), not
// tensor-core throughput.
#include "../../include/pddp_b200.h"
#include "profile.h"
#include <cuda_runtime.h>
#include <stdint.h>

int pddp_capi_fail(int code, const char* msg);
int pddp_capi_cuda(cudaError_t e, const char* what);

namespace pddp {
namespace train {

constexpr int ROWS = 4;        // mini-batch rows per CTA
constexpr int THREADS = 256;   // >= hidden widths
constexpr int MAX_K0 = 16, MAX_OUT = 16;

template <class T>
struct Args {
    int K0, H0, H1, D, OUT, n_data, batch, dropout;
    T lr, beta1, beta2, eps, reg_scale, temperature, reg0, reg1, rate0, rate1;
    uint64_t seed;
    const T* X; const T* dX; const T* X_mean; const T* X_std_inv; const T* dX_mean; const T* dX_std;
    const int32_t* batch_idx; const T* noise;
    T* params; T* grads; T* loss;
    // workspace
    T *W0T, *W1T, *W2T;                  // transposed copies [K0][H0], [H0][H1], [H1][OUT]
    T *A0, *Hid0, *Hid1;                 // activations [batch][K0], [batch][H0], [batch][H1]
    T *DP0, *DP1, *DOUT;                 // deltas w.r.t. the pre-activations / outputs
    T *DL0, *DL1, *LROW;                 // per-row d/dlogit_p and loss
    T *m, *v, *vmax;                     // Adam state [n_params]
    int n_params;
};

// parameter offsets inside the flat vector [W0 | b0 | W1 | b1 | W2 | b2 | logit_p0 | logit_p1]
struct Offsets { int W0, b0, W1, b1, W2, b2, lp0, lp1, end; };
__host__ __device__ inline Offsets offsets(int K0, int H0, int H1, int OUT) {
    Offsets o;
    o.W0 = 0; o.b0 = o.W0 + H0 * K0; o.W1 = o.b0 + H0; o.b1 = o.W1 + H1 * H0; o.W2 = o.b1 + H1;
    o.b2 = o.W2 + OUT * H1; o.lp0 = o.b2 + OUT; o.lp1 = o.lp0 + 1; o.end = o.lp1 + 1;
    return o;
}

// counter-based uniform in (0, 1): splitmix64 of (seed, step, row, unit); 24 bits, never exactly 0 or 1
__device__ __forceinline__ float uniform01(uint64_t seed, int it, int row, int unit) {
    uint64_t x = seed + 0x9E3779B97F4A7C15ull * ((((uint64_t)it << 20) + (uint64_t)row) * 1024ull + (uint64_t)unit + 1ull);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return ((float)(x >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

template <class T> __device__ __forceinline__ T sigmoid(T x) { return T(1) / (T(1) + exp(-x)); }

// dropout mask of one (row, unit) and d mask / d logit_p   (modules.py:540-548 concrete; 462-483 Bernoulli)
template <class T>
__device__ __forceinline__ void mask_of(const Args<T>& a, T r, T lp, T keep_p, T& m, T& dm_dlp) {
    if (a.dropout == 0) {
        m = sigmoid((lp + log(r) - log(T(1) - r)) / a.temperature);
        dm_dlp = m * (T(1) - m) / a.temperature;
    } else {
        m = r < keep_p ? T(1) : T(0);        // torch.bernoulli(p): 1 with probability p
        dm_dlp = T(0);
    }
}

template <class T>
__global__ void __launch_bounds__(THREADS) train_rows_kernel(const Args<T> a, int it) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_a0 = reinterpret_cast<T*>(smem_raw);                 // [ROWS][K0]
    T* s_h0 = s_a0 + ROWS * MAX_K0;                           // [ROWS][H0]
    T* s_h1 = s_h0 + ROWS * THREADS;                          // [ROWS][H1]
    T* s_d1 = s_h1 + ROWS * THREADS;                          // [ROWS][H1]  dpre1
    T* s_out = s_d1 + ROWS * THREADS;                         // [ROWS][OUT] outputs, then dout
    T* s_red = s_out + ROWS * MAX_OUT;                        // [ROWS][THREADS / 32] reduction scratch
    __shared__ int s_idx[ROWS];
    __shared__ int s_nvalid;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * ROWS;
    const Offsets o = offsets(a.K0, a.H0, a.H1, a.OUT);
    const int32_t* idx = a.batch_idx + (int64_t)it * a.batch;
    if (tid < ROWS) s_idx[tid] = row0 + tid < a.batch ? idx[row0 + tid] : -1;
    if (tid == 0) s_nvalid = 0;
    __syncthreads();
    {   // rows of this step that hold data (the last batch of an epoch is partial): loss.mean() divides by it
        int c = 0;
        for (int r = tid; r < a.batch; r += THREADS) c += idx[r] >= 0;
        for (int s = 16; s; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
        if (lane == 0 && c) atomicAdd(&s_nvalid, c);
    }
    for (int e = tid; e < ROWS * a.K0; e += THREADS) {        // normalised inputs (modules.py:122-123)
        const int r = e / a.K0, k = e - r * a.K0;
        T v = T(0);
        if (s_idx[r] >= 0) {
            v = a.X[(int64_t)s_idx[r] * a.K0 + k];
            if (a.X_mean) v = (v - a.X_mean[k]) * a.X_std_inv[k];
        }
        s_a0[r * MAX_K0 + k] = v;
        if (row0 + r < a.batch) a.A0[(int64_t)(row0 + r) * a.K0 + k] = v;
    }
    __syncthreads();
    const T inv_n = T(1) / (T)(s_nvalid > 0 ? s_nvalid : 1);
    const T lp0 = a.params[o.lp0], lp1 = a.params[o.lp1];
    const T keep0 = T(1) - a.rate0, keep1 = T(1) - a.rate1;          // Bernoulli keep-probabilities (dropout == 1)
    const int HH = a.H0 + a.H1;
    auto draw = [&](int r, int unit) -> T {
        if (a.noise) return a.noise[((int64_t)it * a.batch + row0 + r) * HH + unit];
        return (T)uniform01(a.seed, it, row0 + r, unit);
    };

    // ---- layer 0: pre = W0 a + b0, z = pre * mask, h = relu(z) -----------------------------------------
    T pre0[ROWS], m0[ROWS], g0[ROWS];                         // kept in registers for the backward chain
    if (tid < a.H0) {
        T acc[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = a.params[o.b0 + tid];
        for (int k = 0; k < a.K0; ++k) {
            const T w = a.W0T[k * a.H0 + tid];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] += w * s_a0[r * MAX_K0 + k];
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const bool live = s_idx[r] >= 0;
            pre0[r] = acc[r];
            m0[r] = T(0); g0[r] = T(0);
            if (live) mask_of(a, draw(r, tid), lp0, keep0, m0[r], g0[r]);
            const T z = acc[r] * m0[r];
            s_h0[r * THREADS + tid] = z > T(0) ? z : T(0);
        }
    }
    __syncthreads();
    // ---- layer 1 ---------------------------------------------------------------------------------------
    T pre1[ROWS], m1[ROWS], g1[ROWS];
    if (tid < a.H1) {
        T acc[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = a.params[o.b1 + tid];
        for (int k = 0; k < a.H0; ++k) {
            const T w = a.W1T[k * a.H1 + tid];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] += w * s_h0[r * THREADS + k];
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const bool live = s_idx[r] >= 0;
            pre1[r] = acc[r];
            m1[r] = T(0); g1[r] = T(0);
            if (live) mask_of(a, draw(r, a.H0 + tid), lp1, keep1, m1[r], g1[r]);
            const T z = acc[r] * m1[r];
            s_h1[r * THREADS + tid] = z > T(0) ? z : T(0);
        }
    }
    __syncthreads();
    // ---- output layer: one warp per (row, output) pair, lane-strided dot product -------------------------
    for (int p = warp; p < ROWS * a.OUT; p += THREADS / 32) {
        const int r = p / a.OUT, oo = p - r * a.OUT;
        T s = T(0);
        for (int k = lane; k < a.H1; k += 32) s += a.params[o.W2 + oo * a.H1 + k] * s_h1[r * THREADS + k];
        for (int sh = 16; sh; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
        if (lane == 0) s_out[r * MAX_OUT + oo] = s + a.params[o.b2 + oo];
    }
    __syncthreads();
    // ---- loss and d loss / d out per row (losses.py:20-38, modules.py:126-129, 187-191) -------------------
    if (tid < ROWS) {
        const int r = tid;
        T l = T(0);
        if (s_idx[r] >= 0) {
            T quad = T(0), logs = T(0);
            for (int d = 0; d < a.D; ++d) {
                const T sc = a.dX_std ? a.dX_std[d] : T(1), sh = a.dX_mean ? a.dX_mean[d] : T(0);
                const T mean = s_out[r * MAX_OUT + d] * sc + sh;
                const T log_std = s_out[r * MAX_OUT + a.D + d] + log(sc);
                const T delta = mean - a.dX[(int64_t)s_idx[r] * a.D + d];
                const T inv_std = exp(-log_std);
                const T q = delta * inv_std;
                quad += q * q;
                logs += log_std;
                s_out[r * MAX_OUT + d] = q * inv_std * sc * inv_n;              // d/d out_mean
                s_out[r * MAX_OUT + a.D + d] = (T(1) - q * q) * inv_n;          // d/d out_logstd
            }
            l = T(0.5) * quad + logs + T(0.91893853320467274178);               // 0.5 log(2 pi), once per row
        } else {
            for (int d = 0; d < a.OUT; ++d) s_out[r * MAX_OUT + d] = T(0);
        }
        if (row0 + r < a.batch) a.LROW[row0 + r] = l * inv_n;
    }
    __syncthreads();
    for (int e = tid; e < ROWS * a.OUT; e += THREADS) {
        const int r = e / a.OUT, oo = e - r * a.OUT;
        if (row0 + r < a.batch) a.DOUT[(int64_t)(row0 + r) * a.OUT + oo] = s_out[r * MAX_OUT + oo];
    }
    // ---- backward through layer 1 -------------------------------------------------------------------------
    T dl1[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) dl1[r] = T(0);
    if (tid < a.H1) {
        T dh[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) dh[r] = T(0);
        for (int oo = 0; oo < a.OUT; ++oo) {
            const T w = a.W2T[tid * a.OUT + oo];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) dh[r] += w * s_out[r * MAX_OUT + oo];
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const T dz = pre1[r] * m1[r] > T(0) ? dh[r] : T(0);          // relu'
            const T dp = dz * m1[r];
            dl1[r] = dz * pre1[r] * g1[r];                                // through the concrete mask to logit_p
            s_d1[r * THREADS + tid] = dp;
            if (row0 + r < a.batch) {
                a.DP1[(int64_t)(row0 + r) * a.H1 + tid] = dp;
                a.Hid1[(int64_t)(row0 + r) * a.H1 + tid] = s_h1[r * THREADS + tid];
            }
        }
    }
    __syncthreads();
    // ---- backward through layer 0 (W1 in its own [H1][H0] layout: coalesced over this thread's unit) -------
    T dl0[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) dl0[r] = T(0);
    if (tid < a.H0) {
        T dh[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) dh[r] = T(0);
        for (int j = 0; j < a.H1; ++j) {
            const T w = a.params[o.W1 + j * a.H0 + tid];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) dh[r] += w * s_d1[r * THREADS + j];
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const T dz = pre0[r] * m0[r] > T(0) ? dh[r] : T(0);
            dl0[r] = dz * pre0[r] * g0[r];
            if (row0 + r < a.batch) {
                a.DP0[(int64_t)(row0 + r) * a.H0 + tid] = dz * m0[r];
                a.Hid0[(int64_t)(row0 + r) * a.H0 + tid] = s_h0[r * THREADS + tid];
            }
        }
    }
    // ---- d/dlogit_p: sum over the units of a row ----------------------------------------------------------
#pragma unroll
    for (int which = 0; which < 2; ++which) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            T s = which == 0 ? dl0[r] : dl1[r];
            for (int sh = 16; sh; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
            if (lane == 0) s_red[r * (THREADS / 32) + warp] = s;
        }
        __syncthreads();
        if (tid < ROWS && row0 + tid < a.batch) {
            T s = T(0);
            for (int w = 0; w < THREADS / 32; ++w) s += s_red[tid * (THREADS / 32) + w];
            (which == 0 ? a.DL0 : a.DL1)[row0 + tid] = s;
        }
        __syncthreads();
    }
}

// one thread per parameter element: gradient (dot product over the batch rows + regulariser), Adam(amsgrad)
template <class T>
__global__ void __launch_bounds__(256) train_update_kernel(const Args<T> a, int it) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const Offsets o = offsets(a.K0, a.H0, a.H1, a.OUT);
    T g = T(0), regv = T(0);
    T w = T(0);
    if (e < o.end) {
        w = a.params[e];
        const T rs = a.reg_scale / (T)a.n_data;                                   // reg_scale * reg / N
        // The regulariser's keep-probability is ALWAYS 1 - rate, for CDropout too: CDropout.regularization sets
        // p.data = sigmoid(logit_p) and then calls BDropout.regularization, whose first statement rebinds
        // self.p = 1 - self.rate (modules.py:443, 526-527) -- the learned logit_p never reaches the regulariser.
        const T p0 = T(1) - a.rate0, p1 = T(1) - a.rate1;
        const int B = a.batch;
        if (e < o.b0) {                                           // W0[j][k]: no dropout before fc_0, no regulariser
            const int j = e / a.K0, k = e - j * a.K0;
            for (int r = 0; r < B; ++r) g += a.DP0[(int64_t)r * a.H0 + j] * a.A0[(int64_t)r * a.K0 + k];
        } else if (e < o.W1) {
            const int j = e - o.b0;
            for (int r = 0; r < B; ++r) g += a.DP0[(int64_t)r * a.H0 + j];
        } else if (e < o.b1) {                                    // W1: regularised by drop_0 (modules.py:757-765)
            const int j = (e - o.W1) / a.H0, k = (e - o.W1) - j * a.H0;
            for (int r = 0; r < B; ++r) g += a.DP1[(int64_t)r * a.H1 + j] * a.Hid0[(int64_t)r * a.H0 + k];
            g += rs * a.reg0 * p0 * T(2) * w;
            regv = rs * a.reg0 * p0 * w * w;
        } else if (e < o.W2) {
            const int j = e - o.b1;
            for (int r = 0; r < B; ++r) g += a.DP1[(int64_t)r * a.H1 + j];
            g += rs * a.reg0 * T(2) * w;
            regv = rs * a.reg0 * w * w;
        } else if (e < o.b2) {                                    // W2 (fc_out): regularised by drop_1
            const int j = (e - o.W2) / a.H1, k = (e - o.W2) - j * a.H1;
            for (int r = 0; r < B; ++r) g += a.DOUT[(int64_t)r * a.OUT + j] * a.Hid1[(int64_t)r * a.H1 + k];
            g += rs * a.reg1 * p1 * T(2) * w;
            regv = rs * a.reg1 * p1 * w * w;
        } else if (e < o.lp0) {
            const int j = e - o.b2;
            for (int r = 0; r < B; ++r) g += a.DOUT[(int64_t)r * a.OUT + j];
            g += rs * a.reg1 * T(2) * w;
            regv = rs * a.reg1 * w * w;
        } else {                                                  // logit_p of a CDropout: likelihood path only --
            const T* dl = e == o.lp0 ? a.DL0 : a.DL1;             // p enters the regulariser as DATA (modules.py:526)
            const T p = e == o.lp0 ? p0 : p1;
            if (a.dropout == 0) {
                for (int r = 0; r < B; ++r) g += dl[r];
                regv = -rs * (-(T(1) - p) * log(T(1) - p) - p * log(p));          // minus the Bernoulli entropy
            }
            for (int r = 0; r < B && e == o.lp0; ++r) regv += a.LROW[r];          // likelihood part of the loss
        }
        a.grads[e] = g;
        const bool frozen = e >= o.lp0 && a.dropout != 0;                         // BDropout has no logit_p
        if (!frozen) {
            // torch.optim.Adam(amsgrad=True): bias-corrected, max of the second-moment estimate
            const T m = a.beta1 * a.m[e] + (T(1) - a.beta1) * g;
            const T v = a.beta2 * a.v[e] + (T(1) - a.beta2) * g * g;
            const T vm = v > a.vmax[e] ? v : a.vmax[e];
            a.m[e] = m; a.v[e] = v; a.vmax[e] = vm;
            const T bc1 = T(1) - pow(a.beta1, (T)(it + 1)), bc2 = T(1) - pow(a.beta2, (T)(it + 1));
            const T denom = sqrt(vm) / sqrt(bc2) + a.eps;
            w -= (a.lr / bc1) * (m / denom);
            a.params[e] = w;
            if (e < o.b0) { const int j = e / a.K0, k = e - j * a.K0; a.W0T[k * a.H0 + j] = w; }
            else if (e >= o.W1 && e < o.b1) { const int j = (e - o.W1) / a.H0, k = (e - o.W1) - j * a.H0; a.W1T[k * a.H1 + j] = w; }
            else if (e >= o.W2 && e < o.b2) { const int j = (e - o.W2) / a.H1, k = (e - o.W2) - j * a.H1; a.W2T[k * a.OUT + j] = w; }
        }
    }
    // loss of this step (reported, not used): likelihood rows + every element's regulariser share
    for (int s = 16; s; s >>= 1) regv += __shfl_xor_sync(0xffffffffu, regv, s);
    if ((threadIdx.x & 31) == 0 && regv != T(0)) atomicAdd(&a.loss[it], regv);
}

template <class T>
__global__ void train_transpose_kernel(const Args<T> a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const Offsets o = offsets(a.K0, a.H0, a.H1, a.OUT);
    if (e >= o.end) return;
    const T w = a.params[e];
    if (e < o.b0) { const int j = e / a.K0, k = e - j * a.K0; a.W0T[k * a.H0 + j] = w; }
    else if (e >= o.W1 && e < o.b1) { const int j = (e - o.W1) / a.H0, k = (e - o.W1) - j * a.H0; a.W1T[k * a.H1 + j] = w; }
    else if (e >= o.W2 && e < o.b2) { const int j = (e - o.W2) / a.H1, k = (e - o.W2) - j * a.H1; a.W2T[k * a.OUT + j] = w; }
}

static int64_t align16(int64_t x) { return (x + 15) / 16 * 16; }

template <class T>
static int64_t carve(const pddp_bnn_train_config* c, unsigned char* base, Args<T>* a) {
    const int OUT = 2 * c->D;
    const Offsets o = offsets(c->K0, c->H0, c->H1, OUT);
    int64_t off = 0;
    auto take = [&](int64_t n) { T* p = base ? reinterpret_cast<T*>(base + off) : nullptr; off += align16(n * (int64_t)sizeof(T)); return p; };
    T* W0T = take((int64_t)c->K0 * c->H0); T* W1T = take((int64_t)c->H0 * c->H1); T* W2T = take((int64_t)c->H1 * OUT);
    T* A0 = take((int64_t)c->batch * c->K0); T* H0 = take((int64_t)c->batch * c->H0); T* H1 = take((int64_t)c->batch * c->H1);
    T* DP0 = take((int64_t)c->batch * c->H0); T* DP1 = take((int64_t)c->batch * c->H1); T* DOUT = take((int64_t)c->batch * OUT);
    T* DL0 = take(c->batch); T* DL1 = take(c->batch); T* LROW = take(c->batch);
    T* m = take(o.end); T* v = take(o.end); T* vmax = take(o.end);
    if (a) {
        a->W0T = W0T; a->W1T = W1T; a->W2T = W2T; a->A0 = A0; a->Hid0 = H0; a->Hid1 = H1; a->DP0 = DP0; a->DP1 = DP1;
        a->DOUT = DOUT; a->DL0 = DL0; a->DL1 = DL1; a->LROW = LROW; a->m = m; a->v = v; a->vmax = vmax;
        a->n_params = o.end;
    }
    return off;
}

template <class T>
static int run(const pddp_bnn_train_config* c, const void* X, const void* dX, const void* X_mean, const void* X_std_inv,
               const void* dX_mean, const void* dX_std, const int32_t* batch_idx, const void* noise, void* params,
               void* grads, void* loss, void* ws, cudaStream_t st) {
    Args<T> a;
    a.K0 = c->K0; a.H0 = c->H0; a.H1 = c->H1; a.D = c->D; a.OUT = 2 * c->D; a.n_data = c->n_data; a.batch = c->batch;
    a.dropout = c->dropout;
    a.lr = (T)c->lr; a.beta1 = (T)c->beta1; a.beta2 = (T)c->beta2; a.eps = (T)c->eps; a.reg_scale = (T)c->reg_scale;
    a.temperature = (T)c->temperature; a.reg0 = (T)c->reg0; a.reg1 = (T)c->reg1; a.rate0 = (T)c->rate0; a.rate1 = (T)c->rate1;
    a.seed = c->seed;
    a.X = (const T*)X; a.dX = (const T*)dX; a.X_mean = (const T*)X_mean; a.X_std_inv = (const T*)X_std_inv;
    a.dX_mean = (const T*)dX_mean; a.dX_std = (const T*)dX_std; a.batch_idx = batch_idx; a.noise = (const T*)noise;
    a.params = (T*)params; a.grads = (T*)grads; a.loss = (T*)loss;
    const int64_t bytes = carve<T>(c, (unsigned char*)ws, &a);
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)bytes, st);                       // Adam state starts at zero
    if (e != cudaSuccess) return pddp_capi_cuda(e, "pddp_bnn_train(memset)");
    e = cudaMemsetAsync(loss, 0, sizeof(T) * (size_t)c->n_iter, st);
    if (e != cudaSuccess) return pddp_capi_cuda(e, "pddp_bnn_train(memset)");
    const int pgrid = (a.n_params + 255) / 256;
    train_transpose_kernel<T><<<pgrid, 256, 0, st>>>(a);
    const int smem = (int)sizeof(T) * (ROWS * MAX_K0 + 3 * ROWS * THREADS + ROWS * MAX_OUT + ROWS * (THREADS / 32));
    const int rgrid = (c->batch + ROWS - 1) / ROWS;
    for (int it = 0; it < c->n_iter; ++it) {
        train_rows_kernel<T><<<rgrid, THREADS, smem, st>>>(a, it);
        train_update_kernel<T><<<pgrid, 256, 0, st>>>(a, it);
    }
    note_launches(1 + 2LL * c->n_iter);
    return pddp_capi_cuda(cudaGetLastError(), "pddp_bnn_train");
}

}  // namespace train
}  // namespace pddp

using namespace pddp;

static int check_cfg(const pddp_bnn_train_config* c) {
    if (!c) return pddp_capi_fail(PDDP_E_BADARG, "pddp_bnn_train: config is NULL");
    if (c->dtype != PDDP_F32 && c->dtype != PDDP_F64) return pddp_capi_fail(PDDP_E_BADARG, "dtype must be PDDP_F32 or PDDP_F64");
    if (c->K0 < 1 || c->K0 > train::MAX_K0 || c->D < 1 || 2 * c->D > train::MAX_OUT)
        return pddp_capi_fail(PDDP_E_UNSUPPORTED, "pddp_bnn_train: input width <= 16 and state size <= 8");
    if (c->H0 < 1 || c->H0 > train::THREADS || c->H1 < 1 || c->H1 > train::THREADS)
        return pddp_capi_fail(PDDP_E_UNSUPPORTED, "pddp_bnn_train: hidden widths in [1, 256]");
    if (c->n_data < 1 || c->batch < 1 || c->n_iter < 0) return pddp_capi_fail(PDDP_E_BADARG, "pddp_bnn_train: bad sizes");
    if (c->dropout != 0 && c->dropout != 1) return pddp_capi_fail(PDDP_E_BADARG, "pddp_bnn_train: dropout 0 (concrete) or 1 (Bernoulli)");
    return 0;
}

extern "C" int64_t pddp_bnn_train_workspace_bytes(const pddp_bnn_train_config* c) {
    if (int e = check_cfg(c)) return e;
    return c->dtype == PDDP_F32 ? train::carve<float>(c, nullptr, nullptr) : train::carve<double>(c, nullptr, nullptr);
}

extern "C" int pddp_bnn_train(const pddp_bnn_train_config* c, const void* X, const void* dX, const void* X_mean,
                              const void* X_std_inv, const void* dX_mean, const void* dX_std, const int32_t* batch_idx,
                              const void* noise, void* params, void* grads, void* loss, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    if (int e = check_cfg(c)) return e;
    if (!X || !dX || !batch_idx || !params || !grads || !loss || !workspace)
        return pddp_capi_fail(PDDP_E_BADARG, "pddp_bnn_train: NULL argument");
    if ((X_mean == nullptr) != (X_std_inv == nullptr) || (dX_mean == nullptr) != (dX_std == nullptr))
        return pddp_capi_fail(PDDP_E_BADARG, "pddp_bnn_train: normalisation buffers come in pairs");
    if (workspace_bytes < pddp_bnn_train_workspace_bytes(c)) return pddp_capi_fail(PDDP_E_BADARG, "pddp_bnn_train: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (c->dtype == PDDP_F32)
        return train::run<float>(c, X, dX, X_mean, X_std_inv, dX_mean, dX_std, batch_idx, noise, params, grads, loss, workspace, st);
    return train::run<double>(c, X, dX, X_mean, X_std_inv, dX_mean, dX_std, batch_idx, noise, params, grads, loss, workspace, st);
}
