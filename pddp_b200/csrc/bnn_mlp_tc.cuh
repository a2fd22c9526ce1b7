// Particle MLP on the 5th-generation tensor cores (tcgen05 + TMEM), fp32 in / fp32 out.
//
// The 200x200 hidden layer -- the only dense contraction on the PDDP hot path -- runs as
// tcgen05.mma.kind::tf32 with the accumulator in tensor memory.  Plain TF32 (10-bit mantissa) moves
// the feedback gains by ~1e-2 relative (measured), so every operand is split into
// hi + lo TF32 parts and three MMAs (hi*hi + lo*hi + hi*lo) recover fp32-class accuracy ("3xTF32").
//
// Persistent kernel, one CTA per SM, 10 warps, warp-specialised:
//   warps 0-3  epilogue : tcgen05.ld of the 128x208 accumulator (one row per thread), bias + dropout
//                         mask + ReLU, tangent rows follow their primal's activation pattern via
//                         __ballot_sync (a particle's 1+T rows sit in one warp), output layer
//                         H1 -> D on the CUDA cores, X' and dX'/d(X,u) stored to global.
//   warps 4-7  producer : layer 0 (K0 <= 9 inputs) on the CUDA cores for one row per thread, one
//                         32-column K-block at a time, hi/lo split, written to shared memory in the
//                         UMMA K-major SWIZZLE_128B layout (the A operand).
//   warp  8    loader   : streams the pre-swizzled hi/lo images of W1 (the B operand, L2 resident)
//                         with cp.async.bulk + mbarrier complete_tx, one K-block per stage.
//   warp  9    issuer   : one thread issues 4 x 3 tcgen05.mma per K-block and tcgen05.commit's the
//                         stage-empty / accumulator-full barriers.
// Two A/B stages and two TMEM accumulator stages (2 x 256 columns) overlap producer, MMA and
// epilogue across K-blocks and tiles.
#pragma once
#include "bnn_mlp_simt.cuh"

namespace pddp {
namespace tc {

constexpr int TILE_M = 128, TILE_N = 208, KBLK = 32, MAX_KB = 7, STAGES = 2;
constexpr int A_PART_BYTES = TILE_M * 128;            // 16 KB  (128 rows x 32 tf32)
constexpr int B_PART_BYTES = TILE_N * 128;            // 26 KB  (208 rows x 32 tf32)
constexpr int A_STAGE_BYTES = 2 * A_PART_BYTES;       // hi | lo
constexpr int B_STAGE_BYTES = 2 * B_PART_BYTES;       // hi | lo
constexpr int THREADS = 320;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) addr>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4 = 1024 B between
// 8-row groups, [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major A and B,
// N>>3 @17, M>>4 @24.
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

// byte offset of element (row, kk) inside a [rows][32 x tf32] K-major SWIZZLE_128B tile
__host__ __device__ __forceinline__ uint32_t swz_offset(int row, int kk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ (row & 7)) & 7) << 4) + (kk & 3) * 4);
}

struct Smem {
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = A_OFF + STAGES * A_STAGE_BYTES;          // 65536
    static constexpr int W0_OFF = B_OFF + STAGES * B_STAGE_BYTES;         // + 106496
    static constexpr int W0_BYTES = TILE_N * 12 * 4;                      // sW0[c][12], k < K0 <= 9 (float4 reads)
    static constexpr int B0_OFF = W0_OFF + W0_BYTES;
    static constexpr int B1_OFF = B0_OFF + TILE_N * 4;
    static constexpr int W2_OFF = B1_OFF + TILE_N * 4;                    // W2s[c][8]
    static constexpr int BAR_OFF = W2_OFF + TILE_N * 8 * 4;
    static constexpr int TOTAL = BAR_OFF + 16 * 8 + 16;
};

// Image of W1 for the loader: [kb][hi|lo][208 rows x 128 B swizzled]; zero padded.
__global__ void bnn_tc_prep_kernel(const float* W1 /*[H1][H0]*/, int H0, int H1, int nkb, float* img) {
    const int total = nkb * TILE_N * KBLK;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kb = i / (TILE_N * KBLK), rem = i - kb * TILE_N * KBLK;
        const int n = rem / KBLK, kk = rem - n * KBLK, k = kb * KBLK + kk;
        const float w = (n < H1 && k < H0) ? W1[(size_t)n * H0 + k] : 0.f;
        const float hi = tf32_rna(w), lo = tf32_rna(w - hi);
        char* base = reinterpret_cast<char*>(img) + (size_t)kb * B_STAGE_BYTES;
        *reinterpret_cast<float*>(base + swz_offset(n, kk)) = hi;
        *reinterpret_cast<float*>(base + B_PART_BYTES + swz_offset(n, kk)) = lo;
    }
}

template <int GEO, bool TAN>
__global__ void __launch_bounds__(THREADS, 1) bnn_mlp_tc_kernel(const BnnMlpArgs<float> a, const float* __restrict__ Bimg, int nkb) {
    typedef Geo<GEO> G;
    constexpr int D = G::D, DA = G::DA, NNA = G::NNA, NANG = G::NANG, K0 = DA + G::NU;
    constexpr int TD = TAN ? D + G::NU : 0, RPP = 1 + TD, PPW = 32 / RPP, NPART = 4 * PPW;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sW0 = reinterpret_cast<float*>(smem + Smem::W0_OFF);
    float* sb0 = reinterpret_cast<float*>(smem + Smem::B0_OFF);
    float* sb1 = reinterpret_cast<float*>(smem + Smem::B1_OFF);
    float* sW2 = reinterpret_cast<float*>(smem + Smem::W2_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BAR_OFF);
    uint64_t *a_full = bars, *a_empty = bars + 2, *b_full = bars + 4, *b_empty = bars + 6, *acc_full = bars + 8,
             *acc_empty = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const BnnNet<float>& n = a.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H0 = n.H0, H1 = n.H1, P = n.P;
    const long long ntiles = (a.total + NPART - 1) / NPART;

    // ---- one-time setup ----
    for (int i = tid; i < 12 * TILE_N; i += THREADS) {
        const int c = i / 12, k = i - c * 12;
        sW0[i] = (k < K0 && c < H0) ? n.W0T[k * H0 + c] : 0.f;
    }
    for (int i = tid; i < TILE_N; i += THREADS) {
        sb0[i] = i < H0 ? n.b0[i] : 0.f;
        sb1[i] = i < H1 ? n.b1[i] : 0.f;
    }
    for (int i = tid; i < TILE_N * 8; i += THREADS) {
        const int c = i >> 3, o = i & 7;
        sW2[i] = (c < H1 && o < D) ? n.W2T[c * D + o] : 0.f;
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&a_full[s], 128);
            mbar_init(&a_empty[s], 1);
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 && warp < 8) {
        // ================= producer: layer 0 -> A operand =================
        const int r = tid - 128, w = r >> 5, ql = lane / RPP, d = lane - ql * RPP;
        const int q = w * PPW + ql;
        uint32_t kcount = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long g = tile * NPART + q;
            const bool valid = ql < PPW && g < a.total;
            float ap[K0], da[K0];
#pragma unroll
            for (int k = 0; k < K0; ++k) { ap[k] = 0.f; da[k] = 0.f; }
            const float* m0 = n.mask0;
            if (valid) {
                float x[D], in[K0], sc[K0];
#pragma unroll
                for (int i = 0; i < D; ++i) x[i] = a.X[g * D + i];
#pragma unroll
                for (int k = 0; k < K0; ++k) sc[k] = n.X_std_inv ? n.X_std_inv[k] : 1.f;
#pragma unroll
                for (int i = 0; i < NNA; ++i) in[i] = x[G::nonang(i)];
#pragma unroll
                for (int i = 0; i < NANG; ++i) { in[NNA + 2 * i] = sinf(x[G::ang(i)]); in[NNA + 2 * i + 1] = cosf(x[G::ang(i)]); }
                in[DA] = a.u[g / P];
#pragma unroll
                for (int k = 0; k < K0; ++k) ap[k] = (in[k] - (n.X_mean ? n.X_mean[k] : 0.f)) * sc[k];
                if (TAN && d > 0) {
                    const int dir = d - 1;
#pragma unroll
                    for (int i = 0; i < NNA; ++i) if (dir == G::nonang(i)) da[i] = sc[i];
#pragma unroll
                    for (int i = 0; i < NANG; ++i) if (dir == G::ang(i)) {
                        da[NNA + 2 * i] = in[NNA + 2 * i + 1] * sc[NNA + 2 * i];
                        da[NNA + 2 * i + 1] = -in[NNA + 2 * i] * sc[NNA + 2 * i + 1];
                    }
                    if (dir == D) da[DA] = sc[DA];
                }
                m0 = n.mask0 + (size_t)(g % P) * H0;
            }
            for (int kb = 0; kb < nkb; ++kb, ++kcount) {
                const int s = kcount & 1;
                mbar_wait(&a_empty[s], ((kcount >> 1) & 1) ^ 1);
                unsigned char* Ahi = smem + Smem::A_OFF + s * A_STAGE_BYTES;
                unsigned char* Alo = Ahi + A_PART_BYTES;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float hi[4], lo[4], mk4[4];
                    // this row's dropout-mask values for 4 columns: one LDG.128 (H0 % 4 == 0 checked on the host)
                    if (valid && kb * KBLK + j * 4 < H0)
                        *reinterpret_cast<float4*>(mk4) = __ldg(reinterpret_cast<const float4*>(m0 + kb * KBLK + j * 4));
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = kb * KBLK + j * 4 + e;
                        float v = 0.f;
                        if (valid && c < H0) {
                            float wk[12];
#pragma unroll
                            for (int k4 = 0; k4 < (K0 + 3) / 4; ++k4)
                                *reinterpret_cast<float4*>(wk + 4 * k4) = *reinterpret_cast<const float4*>(sW0 + c * 12 + 4 * k4);
                            float pre = sb0[c];
#pragma unroll
                            for (int k = 0; k < K0; ++k) pre += ap[k] * wk[k];
                            const float mk = mk4[e];
                            const float pm = pre * mk;
                            if (!TAN || d == 0) v = pm > 0.f ? pm : 0.f;
                            else if (pm > 0.f) {
                                float dp = 0.f;
#pragma unroll
                                for (int k = 0; k < K0; ++k) dp += da[k] * wk[k];
                                v = dp * mk;
                            }
                        }
                        hi[e] = tf32_rna(v);
                        lo[e] = tf32_rna(v - hi[e]);
                    }
                    const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((j ^ (r & 7)) & 7) << 4));
                    *reinterpret_cast<float4*>(Ahi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(Alo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_async_smem();
                mbar_arrive(&a_full[s]);
            }
        }
    } else if (warp == 8) {
        // ================= loader: W1 hi/lo images -> B operand =================
        if (lane == 0) {
            uint32_t kcount = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int kb = 0; kb < nkb; ++kb, ++kcount) {
                    const int s = kcount & 1;
                    mbar_wait(&b_empty[s], ((kcount >> 1) & 1) ^ 1);
                    mbar_expect_tx(&b_full[s], B_STAGE_BYTES);
                    bulk_g2s(smem + Smem::B_OFF + s * B_STAGE_BYTES,
                             reinterpret_cast<const char*>(Bimg) + (size_t)kb * B_STAGE_BYTES, B_STAGE_BYTES, &b_full[s]);
                }
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t kcount = 0, it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                mbar_wait(&acc_empty[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
                for (int kb = 0; kb < nkb; ++kb, ++kcount) {
                    const int s = kcount & 1;
                    const uint32_t ph = (kcount >> 1) & 1;
                    mbar_wait(&a_full[s], ph);
                    mbar_wait(&b_full[s], ph);
                    tc_fence_after();
                    const uint64_t ahi = make_desc(smem_u32(smem + Smem::A_OFF + s * A_STAGE_BYTES));
                    const uint64_t alo = make_desc(smem_u32(smem + Smem::A_OFF + s * A_STAGE_BYTES + A_PART_BYTES));
                    const uint64_t bhi = make_desc(smem_u32(smem + Smem::B_OFF + s * B_STAGE_BYTES));
                    const uint64_t blo = make_desc(smem_u32(smem + Smem::B_OFF + s * B_STAGE_BYTES + B_PART_BYTES));
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {      // UMMA_K = 8 tf32 = 32 B -> +2 in the >>4 address field
                        const uint64_t o = (uint64_t)(k4 * 2);
                        tc_mma_tf32(d_tmem, ahi + o, bhi + o, IDESC, (kb | k4) != 0);
                        tc_mma_tf32(d_tmem, alo + o, bhi + o, IDESC, 1);
                        tc_mma_tf32(d_tmem, ahi + o, blo + o, IDESC, 1);
                    }
                    tc_commit(&a_empty[s]);
                    tc_commit(&b_empty[s]);
                }
                tc_commit(&acc_full[as]);
            }
        }
    } else {
        // ================= epilogue: warps 0-3, one accumulator row per thread =================
        const int w = warp, ql = lane / RPP, d = lane - ql * RPP, q = w * PPW + ql;
        const int primal_lane = ql * RPP;
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const long long g = tile * NPART + q;
            const bool valid = ql < PPW && g < a.total;
            const float* m1 = n.mask1 + (valid ? (size_t)(g % P) * H1 : 0);
            mbar_wait(&acc_full[as], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(as * 256);
            float y[D];
#pragma unroll
            for (int o = 0; o < D; ++o) y[o] = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < TILE_N; c0 += 16) {
                float acc[16], mk16[16];
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    if (valid && c0 + 4 * e4 < H1)
                        *reinterpret_cast<float4*>(mk16 + 4 * e4) = __ldg(reinterpret_cast<const float4*>(m1 + c0 + 4 * e4));
                    else
                        mk16[4 * e4] = mk16[4 * e4 + 1] = mk16[4 * e4 + 2] = mk16[4 * e4 + 3] = 0.f;
                }
                tc_ld16(taddr + c0, acc);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int c = c0 + e;
                    const float mk = mk16[e];
                    const float pm = (acc[e] + sb1[c]) * mk;          // meaningful on primal rows
                    float v;
                    if (TAN) {
                        const unsigned on = __ballot_sync(0xffffffffu, pm > 0.f);
                        const bool act = (on >> primal_lane) & 1u;
                        v = d == 0 ? (pm > 0.f ? pm : 0.f) : (act ? acc[e] * mk : 0.f);
                    } else {
                        v = pm > 0.f ? pm : 0.f;
                    }
                    float w2[8];
#pragma unroll
                    for (int o4 = 0; o4 < (D + 3) / 4; ++o4)
                        *reinterpret_cast<float4*>(w2 + 4 * o4) = *reinterpret_cast<const float4*>(sW2 + c * 8 + 4 * o4);
#pragma unroll
                    for (int o = 0; o < D; ++o) y[o] += v * w2[o];
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[as]);
            if (valid) {
#pragma unroll
                for (int o = 0; o < D; ++o) {
                    const float sd = n.dX_std ? n.dX_std[o] : 1.f, mn = n.dX_mean ? n.dX_mean[o] : 0.f;
                    if (!TAN || d == 0) a.Xn[g * D + o] = a.X[g * D + o] + ((y[o] + n.b2[o]) * sd + mn);
                    else a.Jp[(g * D + o) * TD + (d - 1)] = ((d - 1) == o ? 1.f : 0.f) + y[o] * sd;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

}  // namespace tc
}  // namespace pddp
