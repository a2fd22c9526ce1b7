// Particle MLP on the tensor cores, CTA-PAIR version of bnn_mlp_tc.cuh (tcgen05 cta_group::2).
//
// Same roles, tiles, images and arithmetic as bnn_mlp_tc_kernel; what changes is who feeds the MMAs.  The
// single-CTA kernel is bound by the shared-memory data pipe (profiles/r1_summary.md: 98 % busy with MMA operand
// fetch + LDS/STS + bulk-copy fill).  Here two CTAs of a cluster (one TPC) work on one 256-row tile per track:
// every MMA is M = 256 and is issued by the leader CTA; each CTA supplies its own 128 rows of A and only
// HALF of B (104 of the 208 W1 rows, 16 of the 32 layer-0 rows) -- the tensor cores exchange the B halves --
// so per CTA the B operand fetch and the bulk-copy fill are halved (~20 % fewer shared-memory wavefronts per tile).
//   * consumer -> issuer barriers (a0_full, acc0_empty, a1_full, acc1_empty) live in the leader and count the
//     producer threads of BOTH CTAs (the peer arrives remotely, mapa + mbarrier.arrive.release.cluster);
//   * issuer -> consumer signals are tcgen05.commit.cta_group::2 ... multicast::cluster to the barrier at the
//     same offset in both CTAs (acc0_full, a1_empty, b_empty, acc1_full);
//   * each CTA bulk-copies its half of every W1 K-block / layer-0 image; the peer's otherwise idle issuer warps
//     relay "landed" to the leader (b_peer, w0_peer).
#pragma once
#include "bnn_mlp_tc.cuh"

namespace pddp {
namespace tc {

constexpr int B_HALF = B_STAGE / 2;       // 104 rows x 64 B

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void arrive_leader(uint64_t* bar, uint32_t rank) {
    if (rank == 0) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    else mbar_arrive_remote(bar, 0);
}
// one arrival per warp: a remote mbarrier.arrive is a DSMEM transaction (~200 cycles, low throughput); 128 of them per
// barrier phase made the pair kernel 2x SLOWER than the single-CTA one
__device__ __forceinline__ void warp_arrive_leader(uint64_t* bar, uint32_t rank, int lane) {
    __syncwarp();
    if (lane == 0) arrive_leader(bar, rank);
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra.uni WAIT_DONE_C;\n\t"
        "bra.uni WAIT_LOOP_C;\n\t"
        "WAIT_DONE_C:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(MBAR_SUSPEND_NS) : "memory");
}
__device__ __forceinline__ bool mbar_test_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}

// waits without a suspend-time hint (experiment: are remotely completed phases slow to wake a suspended waiter?)
__device__ __forceinline__ void mbar_wait_fast(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_F:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE_F;\n\t"
        "bra.uni WAIT_LOOP_F;\n\t"
        "WAIT_DONE_F:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_fast(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_G:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE_G;\n\t"
        "bra.uni WAIT_LOOP_G;\n\t"
        "WAIT_DONE_G:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int K0P, int DP>
struct Cfg2 {
    // RES: this CTA's half of the whole W1 image (13 x 6.5 KB) stays RESIDENT in shared memory: no W1 stream, no loader
    // loop, no b_full / b_empty traffic, and the two tracks no longer advance in lockstep through a shared ring.  Half an
    // image fits only because cta_group::2 lets each CTA hold half of B; with the 16-wide layer-0 operand (double
    // cartpole) it does not, and that geometry keeps the ring.  (Measured: -3 % on the rollout launch.  Making the MMA
    // phases of the two tracks mutually exclusive on top of this changed nothing: see profiles/r1_summary.md section 5.)
    static constexpr bool RES = K0P == 8;
    static constexpr int NB = RES ? MAX_NKB : 6;   // resident K-blocks / W1 K-block stages (half blocks: 6.5 KB each)
    static constexpr int NS = 6;                   // layer-1 A-operand slots per track
    static constexpr int ROWB0 = K0P * 4;
    static constexpr int A0_BYTES = TILE_M * ROWB0;
    static constexpr int W0_CHUNK_PART = N0 * ROWB0, W0_CHUNK = 2 * W0_CHUNK_PART, W0_BYTES = MAX_NCH * W0_CHUNK;
    static constexpr int W0_PAD = 1024;            // >= 16 rows of the layer-0 image (the peer's early landing)
    static constexpr int W2_BYTES = TILE_N * DP * 4;
    static constexpr int B_OFF = 0;
    static constexpr int A1_OFF = B_OFF + NB * B_HALF;
    static constexpr int A0_OFF = A1_OFF + 2 * NS * A1_SLOT;
    static constexpr int W0_OFF = A0_OFF + 2 * A0_BYTES + W0_PAD;
    static constexpr int W2_OFF = W0_OFF + 2 * W0_BYTES;
    static constexpr int BAR_OFF = W2_OFF + 2 * W2_BYTES;
    static constexpr int NBARS = 2 * NB + 4 * NS + 14 + NB + 2;
    static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;
    static constexpr int ALIGN_PAD = 512;
    static_assert(16 * ROWB0 <= W0_PAD && A1_OFF % 512 == 0 && W0_OFF % 512 == 0, "layout");
};

template <int GEO, bool TAN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
bnn_mlp_tc2_kernel(const BnnMlpArgs<float> a, const Images im, int S, int tiles_p, int nkb) {
    typedef Geo<GEO> G;
    constexpr int D = G::D, DA = G::DA, NNA = G::NNA, NANG = G::NANG, K0 = DA + G::NU;
    constexpr int K0P = K0 + 1 <= 8 ? 8 : 16, DP = D <= 4 ? 4 : 8;
    typedef Cfg2<K0P, DP> C;
    constexpr int NB = C::NB, NS = C::NS, ROWB0 = C::ROWB0;
    constexpr bool RES = C::RES;
    constexpr int TD = TAN ? D + G::NU : 0, RPP = 1 + TD;     // passes per super-tile: primal + one per tangent direction
    constexpr uint32_t IDESC1 = idesc_f16(2 * TILE_M, TILE_N), IDESC0 = idesc_f16(2 * TILE_M, N0);   // M = 256 over the CTA pair

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + C::ALIGN_PAD - 1) & ~(uintptr_t)(C::ALIGN_PAD - 1));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* b_full = bars;                 // [NB]
    uint64_t* b_empty = b_full + NB;         // [NB]
    uint64_t* a1_full = b_empty + NB;        // [2][NS]
    uint64_t* a1_empty = a1_full + 2 * NS;   // [2][NS]
    uint64_t* acc0_full = a1_empty + 2 * NS; // [2]
    uint64_t* acc0_empty = acc0_full + 2;
    uint64_t* a0_full = acc0_empty + 2;
    uint64_t* acc1_full = a0_full + 2;
    uint64_t* acc1_empty = acc1_full + 2;
    uint64_t* w0_full = acc1_empty + 2;
    uint64_t* w2_full = w0_full + 2;
    uint64_t* b_peer = w2_full + 2;          // [NB]  leader only: the peer CTA's half of the W1 K-block has landed
    uint64_t* w0_peer = b_peer + NB;         // [2]   leader only: the peer CTA's half of the layer-0 image has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBARS);
    const uint32_t rank = cluster_ctarank();  // 0 = leader (issues every MMA of the pair), 1 = peer

    const BnnNet<float>& n = a.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P = n.P;

    // ---- tile schedule.  A super-tile is 128 items (rollout: (problem, alpha) pairs; linearise:
    // problems) of one particle; it is RPP consecutive MMA tiles ("passes") on one track: the primal
    // rows, then one pass per tangent direction -- so a thread sees the primal and the tangent
    // pre-activations of the same (item, hidden unit) and the ReLU gate never crosses lanes.
    // Each CTA owns a contiguous range of super-tiles; track t takes every other one.  Track 1 is
    // `skew` K-blocks behind track 0 in the W1 stream.
    const long long NT = (long long)P * tiles_p;
    const long long ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;       // clusters of two CTAs share a schedule
    const long long T0 = NT * cid / ncl, T1 = NT * (cid + 1) / ncl;
    const int cnt = (int)(T1 - T0);
    const int ntl[2] = {((cnt + 1) / 2) * RPP, (cnt / 2) * RPP};     // MMA tiles per track
    const int skew = (nkb / 2) & ~1;
    const int nch = (nkb + 1) / 2;
    const uint32_t w0_bytes = (uint32_t)nch * C::W0_CHUNK;
    const int len0 = ntl[0] * nkb, len1 = ntl[1] ? skew + ntl[1] * nkb : 0;
    const int nblk = len0 > len1 ? len0 : len1;          // W1 K-blocks this CTA streams

    if (tid == 0) {
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 2); }   // both tracks release a W1 stage
        for (int s = 0; s < 2 * NS; ++s) { mbar_init(&a1_full[s], 8); mbar_init(&a1_empty[s], 1); }   // one arrival per producer warp of both CTAs
        for (int s = 0; s < NB; ++s) mbar_init(&b_peer[s], 1);
        for (int t = 0; t < 2; ++t) {
            mbar_init(&acc0_full[t], 1);
            mbar_init(&acc0_empty[t], 8);
            mbar_init(&a0_full[t], 8);
            mbar_init(&acc1_full[t], 1);
            mbar_init(&acc1_empty[t], 8);
            mbar_init(&w0_peer[t], 1);
            mbar_init(&w0_full[t], 1);
            mbar_init(&w2_full[t], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync();                            // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // peer CTA: report the arrival of track t's layer-0 half image to the leader, once per particle change
    auto w0_relay = [&](int t) {
        uint32_t loads = 0;
        int curp = -1;
        for (int ks = 0; ks * RPP < ntl[t]; ++ks) {
            const int p = (int)((T0 + 2 * ks + t) / tiles_p);
            if (p == curp) continue;
            curp = p;
            mbar_wait(&w0_full[t], loads & 1);
            ++loads;
            mbar_arrive_remote(&w0_peer[t], 0);
        }
    };
    if (warp == 18) {
        // ================= loader: W1 K-blocks, shared by both tracks =================
        if (lane == 0 && RES) {
            mbar_expect_tx(&b_full[0], (uint32_t)nkb * B_HALF);       // the whole half image, once
            for (int kb = 0; kb < nkb; ++kb)
                bulk_g2s(smem + C::B_OFF + kb * B_HALF, im.W1img + (size_t)kb * B_STAGE + rank * B_HALF, B_HALF, &b_full[0]);
        } else if (lane == 0) {
            uint32_t s = 0, ph = 1, kb = 0;
            for (int nb = 0; nb < nblk; ++nb) {
                mbar_wait(&b_empty[s], ph);
                mbar_expect_tx(&b_full[s], B_HALF);    // rows [104 rank, 104 rank + 104) of the K-block: this CTA's half of B
                bulk_g2s(smem + C::B_OFF + s * B_HALF, im.W1img + (size_t)kb * B_STAGE + rank * B_HALF, B_HALF, &b_full[s]);
                if (++s == NB) { s = 0; ph ^= 1; }
                if (++kb == (uint32_t)nkb) kb = 0;
            }
        }
    } else if (warp == 19) {
        // ================= layer-0 issuer (one thread, both tracks, polling): chunk c+1 of a track is
        // issued the moment its mid-stage has drained chunk c, independent of where layer 1 stands ====
        if (lane == 0 && rank != 0) {
            w0_relay(1);                       // peer: tell the leader when track 1's half image has landed
        } else if (lane == 0) {
            uint32_t l0cnt[2] = {0, 0}, w0loads[2] = {0, 0};
            int curp[2] = {-1, -1}, k[2] = {0, 0}, pos[2] = {0, 0};
            bool fresh[2] = {true, true};          // next chunk is the first of its tile
            while (k[0] < ntl[0] || k[1] < ntl[1]) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (k[t] >= ntl[t]) continue;
                    if (fresh[t]) {
                        const int p = (int)((T0 + 2 * (k[t] / RPP) + t) / tiles_p);
                        if (p != curp[t]) {
                            if (!mbar_test(&w0_full[t], w0loads[t] & 1) || !mbar_test_cluster(&w0_peer[t], w0loads[t] & 1)) continue;
                            ++w0loads[t];
                            curp[t] = p;
                        }
                        if (!mbar_test_cluster(&a0_full[t], (uint32_t)k[t] & 1)) continue;
                        fresh[t] = false;
                    }
                    if (!mbar_test_cluster(&acc0_empty[t], (l0cnt[t] & 1) ^ 1)) continue;
                    tc_fence_after();
                    int kb = t * skew + pos[t];
                    if (kb >= nkb) kb -= nkb;
                    const int nk = kb + 1 < nkb ? 2 : 1, j = kb >> 1;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(t * TM_TRACK + TM_ACC0);
                    const uint32_t a0 = smem_u32(smem + C::A0_OFF + t * C::A0_BYTES);
                    const uint32_t b0 = smem_u32(smem + C::W0_OFF + t * C::W0_BYTES + j * C::W0_CHUNK);
                    const uint64_t ad = make_desc<ROWB0>(a0);
                    const uint64_t bx = make_desc<ROWB0>(b0), by = make_desc<ROWB0>(b0 + C::W0_CHUNK_PART);
                    // one K-step = 32 B of a row = 16 fp16: K0P = 8 -> [a0 | a1] in one step; K0P = 16 -> a0, then a1
                    tc_mma2_f16(d_tmem, ad, bx, IDESC0, 0);                          // a0*b0 (+ a1*b0)
                    if (K0P == 16) tc_mma2_f16(d_tmem, ad + 2, bx + 2, IDESC0, 1);   // a1*b0
                    tc_mma2_f16(d_tmem, ad, by, IDESC0, 1);                          // a0*b1
                    tc_commit2(&acc0_full[t]);
                    ++l0cnt[t];
                    pos[t] += nk;
                    if (pos[t] >= nkb) { pos[t] = 0; ++k[t]; fresh[t] = true; }
                }
            }
        }
    } else if (warp >= 16) {
        // ================= layer-1 issuers: one thread per track (a track that waits for its epilogue
        // or its mid-stage does not hold the other one up; they drift at most NB W1 stages apart) ====
        // The whole warp runs the loop with warp-uniform state (so the compiler keeps it in uniform
        // registers and feeds UTCHMMA without per-operand R2UR / elect loops: the single-lane version
        // spent ~1000 cycles of dependent scalar code per step for 318 cycles of MMA); one elected lane
        // issues.  Descriptors advance by adding to their low word (address field, 16-byte units).
        if (rank != 0) {
            // peer CTA: these two warps only relay "my half has landed" to the leader (warp 16: W1 K-blocks,
            // warp 17: track 0's layer-0 image)
            if (lane == 0) {
                if (warp == 16 && RES) {
                    mbar_wait(&b_full[0], 0);
                    mbar_arrive_remote(&b_peer[0], 0);
                } else if (warp == 16) {
                    uint32_t s = 0, ph = 0;
                    for (int nb = 0; nb < nblk; ++nb) {
                        mbar_wait(&b_full[s], ph);
                        mbar_arrive_remote(&b_peer[s], 0);
                        if (++s == NB) { s = 0; ph ^= 1; }
                    }
                } else {
                    w0_relay(0);
                }
            }
        } else {
            const int t = __shfl_sync(0xffffffffu, warp, 0) - 16;
            const bool leader = elect_one();
            const int ntiles = t == 0 ? ntl[0] : ntl[1], first_blk = t * skew;
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * TM_TRACK + TM_ACC1);
            const uint64_t desc_hi = make_desc<64>(0) & 0xFFFFFFFF00000000ull;
            const uint32_t a_lo0 = (uint32_t)make_desc<64>(smem_u32(smem + C::A1_OFF + t * NS * A1_SLOT));
            const uint32_t b_lo0 = (uint32_t)make_desc<64>(smem_u32(smem + C::B_OFF));   // this CTA's half; the peer's sits at the same offset
            uint64_t* const a1f = a1_full + t * NS;
            uint64_t* const a1e = a1_empty + t * NS;
            uint32_t s = 0, bph = 0, slot = 0, sph = 0;
            int k = 0, pos = 0;
            if (RES) {
                mbar_wait(&b_full[0], 0);
                mbar_wait_cluster(&b_peer[0], 0);
                for (k = 0; k < ntiles; ++k) {
                    mbar_wait_cluster(&acc1_empty[t], ((uint32_t)k & 1) ^ 1);
                    for (pos = 0; pos < nkb; ++pos) {
                        int kb = first_blk + pos;
                        if (kb >= nkb) kb -= nkb;
                        mbar_wait_cluster(&a1f[slot], sph);
                        tc_fence_after();
                        if (leader) {
                            const uint64_t ad = desc_hi | (uint64_t)(a_lo0 + slot * (A1_SLOT >> 4));
                            const uint64_t bd = desc_hi | (uint64_t)(b_lo0 + (uint32_t)kb * (B_HALF >> 4));
                            tc_mma2_f16(d_tmem, ad, bd, IDESC1, pos != 0);     // a0 * b0
                            tc_mma2_f16(d_tmem, ad, bd + 2, IDESC1, 1);        // a0 * b1
                            tc_mma2_f16(d_tmem, ad + 2, bd, IDESC1, 1);        // a1 * b0
                            tc_commit2(&a1e[slot]);
                            if (pos == nkb - 1) tc_commit2(&acc1_full[t]);
                        }
                        __syncwarp();
                        if (++slot == NS) { slot = 0; sph ^= 1; }
                    }
                }
            } else
            for (int nb = 0; nb < nblk; ++nb) {
                mbar_wait(&b_full[s], bph);
                mbar_wait_cluster(&b_peer[s], bph);
                if (nb < first_blk || k >= ntiles) {
                    if (leader) { mbar_arrive(&b_empty[s]); mbar_arrive_remote(&b_empty[s], 1); }   // this track does not use the block
                } else {
                    if (pos == 0) mbar_wait_cluster(&acc1_empty[t], ((uint32_t)k & 1) ^ 1);
                    mbar_wait_cluster(&a1f[slot], sph);
                    tc_fence_after();
                    if (leader) {
                        const uint64_t ad = desc_hi | (uint64_t)(a_lo0 + slot * (A1_SLOT >> 4));
                        const uint64_t bd = desc_hi | (uint64_t)(b_lo0 + s * (B_HALF >> 4));
                        // one UMMA K-step = 32 B of a row = 16 fp16: [x0 | x1] halves are +2 apart in the >>4 address field
                        tc_mma2_f16(d_tmem, ad, bd, IDESC1, pos != 0);     // a0 * b0
                        tc_mma2_f16(d_tmem, ad, bd + 2, IDESC1, 1);        // a0 * b1
                        tc_mma2_f16(d_tmem, ad + 2, bd, IDESC1, 1);        // a1 * b0
                        tc_commit2(&a1e[slot]);
                        tc_commit2(&b_empty[s]);
                        if (pos == nkb - 1) tc_commit2(&acc1_full[t]);
                    }
                    __syncwarp();
                    if (++slot == NS) { slot = 0; sph ^= 1; }
                    if (++pos == nkb) { pos = 0; ++k; }
                }
                if (++s == NB) { s = 0; bph ^= 1; }
            }
        }
    } else {
        // ================= worker teams: warps 0-7 epilogue (track 0, 1), warps 8-15 mid-stage =================
        const int team = warp >> 2, t = team & 1;
        const int r = tid & 127, w = r >> 5;                  // w = TMEM lane quarter of this warp
        const uint32_t track_taddr = tmem_base + (uint32_t)(t * TM_TRACK);
        const int nt = ntl[t];
        if (nt == 0) goto done;
        if (team >= 2) {
            // ---------------- mid-stage (thread = row r): inputs -> A0, layer-0 accumulator -> A1 ----------------
            const uint32_t A0 = smem_u32(smem + C::A0_OFF + t * C::A0_BYTES);
            const uint32_t lane_taddr = track_taddr + ((uint32_t)(w * 32) << 16);
            const uint32_t row_off = (uint32_t)((r >> 3) * 512 + (r & 7) * 64), row_sw = (uint32_t)((r >> 1) & 3);   // SWIZZLE_64B row
            float inc[K0], inn[K0];            // features [aug(x), u] of this row's item: current / next super-tile
            bool vc = false, vn = false;
            uint32_t gate[MAX_NCH];            // sign bits of the primal pre-activations, one word per chunk (TAN)
#pragma unroll
            for (int j = 0; j < MAX_NCH; ++j) gate[j] = 0;
            auto fetch = [&](int ks) {         // item r of super-tile ks: particle p, item i -> global row i*P + p
                const long long tau = T0 + 2 * ks + t;
                const int p = (int)(tau / tiles_p), l = (int)(tau - (long long)p * tiles_p);
                const int i = l * (2 * TILE_M) + (int)rank * TILE_M + r;    // a pair tile is 256 items, this CTA's half
                vn = i < S;
#pragma unroll
                for (int k = 0; k < K0; ++k) inn[k] = 0.f;
                if (vn) {
                    float x[D];
                    const float* xp = a.X + ((size_t)i * P + p) * D;
                    if (D % 4 == 0) {
#pragma unroll
                        for (int e = 0; e < D; e += 4) *reinterpret_cast<float4*>(x + e) = __ldg(reinterpret_cast<const float4*>(xp + e));
                    } else {
#pragma unroll
                        for (int e = 0; e < D; e += 2) *reinterpret_cast<float2*>(x + e) = __ldg(reinterpret_cast<const float2*>(xp + e));
                    }
#pragma unroll
                    for (int i2 = 0; i2 < NNA; ++i2) inn[i2] = x[G::nonang(i2)];
#pragma unroll
                    for (int i2 = 0; i2 < NANG; ++i2) { inn[NNA + 2 * i2] = sinf(x[G::ang(i2)]); inn[NNA + 2 * i2 + 1] = cosf(x[G::ang(i2)]); }
                    inn[DA] = __ldg(a.u + i);
                }
            };
            auto write_a0 = [&](int d) {       // [norm(aug(x), u), 1] (d == 0) or its tangent along direction d-1
                float row[K0P];
#pragma unroll
                for (int k = 0; k < K0P; ++k) row[k] = 0.f;
                if (vc) {
                    float sc[K0];
#pragma unroll
                    for (int k = 0; k < K0; ++k) sc[k] = n.X_std_inv ? n.X_std_inv[k] : 1.f;
                    if (!TAN || d == 0) {
#pragma unroll
                        for (int k = 0; k < K0; ++k) row[k] = (inc[k] - (n.X_mean ? n.X_mean[k] : 0.f)) * sc[k];
                        row[K0] = 1.f;
                    } else {
                        const int dir = d - 1;
#pragma unroll
                        for (int i2 = 0; i2 < NNA; ++i2) if (dir == G::nonang(i2)) row[i2] = sc[i2];
#pragma unroll
                        for (int i2 = 0; i2 < NANG; ++i2) if (dir == G::ang(i2)) {
                            row[NNA + 2 * i2] = inc[NNA + 2 * i2 + 1] * sc[NNA + 2 * i2];          // d sin = cos
                            row[NNA + 2 * i2 + 1] = -inc[NNA + 2 * i2] * sc[NNA + 2 * i2 + 1];     // d cos = -sin
                        }
                        if (dir == D) row[DA] = sc[DA];
                    }
                }
                uint32_t x0[K0P / 2], x1[K0P / 2];       // fp16 pairs: x0 = fp16(row), x1 = fp16(row - x0)
#pragma unroll
                for (int e = 0; e < K0P / 2; ++e) {
                    x0[e] = pack_f16(row[2 * e], row[2 * e + 1]);
                    const float2 back = unpack_f16(x0[e]);
                    x1[e] = pack_f16(row[2 * e] - back.x, row[2 * e + 1] - back.y);
                }
#pragma unroll
                for (int c = 0; c < K0P / 8; ++c) {      // 16-byte chunks: [x0 ...][x1 ...]
                    sts128(A0 + swz_byte<ROWB0>(r, 16 * c), x0[4 * c], x0[4 * c + 1], x0[4 * c + 2], x0[4 * c + 3]);
                    sts128(A0 + swz_byte<ROWB0>(r, 16 * (K0P / 8 + c)), x1[4 * c], x1[4 * c + 1], x1[4 * c + 2], x1[4 * c + 3]);
                }
                fence_async_smem();
                warp_arrive_leader(&a0_full[t], rank, lane);
            };
            auto load_w0 = [&](int p) {
                mbar_expect_tx(&w0_full[t], w0_bytes);
                // the image holds 32-row parts; a CTA feeds rows [16 rank, 16 rank + 16) of every part, which must sit at the
                // same offset in both CTAs: the peer lands the whole image 16 rows early (its rows 0-15 of a part fall into
                // the unused upper half of the previous part, or into W0_PAD)
                bulk_g2s(smem + C::W0_OFF + t * C::W0_BYTES - rank * (16 * ROWB0), im.W0img + (size_t)p * C::W0_BYTES, w0_bytes, &w0_full[t]);
            };
            int curp = (int)((T0 + t) / tiles_p);
            if (r == 0) load_w0(curp);
            fetch(0);
#pragma unroll
            for (int k = 0; k < K0; ++k) inc[k] = inn[k];
            vc = vn;
            write_a0(0);
            uint32_t ci = 0, slot = 0, sph = 1;
            const int kb0 = t * skew;
            int ks = 0, d = 0;                 // super-tile and pass of the current tile
            for (int k = 0; k < nt; ++k) {
                const bool last_pass = d == RPP - 1;
                if (last_pass && k + 1 < nt) fetch(ks + 1);
                for (int pos = 0; pos < nkb;) {
                    int kb = kb0 + pos;
                    if (kb >= nkb) kb -= nkb;
                    const int nk = kb + 1 < nkb ? 2 : 1, j = kb >> 1;
                    float v[32];
                    mbar_wait(&acc0_full[t], ci & 1);
                    ++ci;
                    tc_fence_after();
                    tc_ld32(lane_taddr + TM_ACC0, v);
                    tc_wait_ld32(v);
                    tc_fence_before();
                    warp_arrive_leader(&acc0_empty[t], rank, lane);
                    if (pos + nk >= nkb && k + 1 < nt) {
                        // every layer-0 MMA of this tile has completed: A0 (and, between super-tiles, the W0 image) is free
                        if (last_pass) {
                            const int pn = (int)((T0 + 2 * (ks + 1) + t) / tiles_p);
                            if (pn != curp) { if (r == 0) load_w0(pn); curp = pn; }
#pragma unroll
                            for (int k2 = 0; k2 < K0; ++k2) inc[k2] = inn[k2];
                            vc = vn;
                            write_a0(0);
                        } else {
                            write_a0(d + 1);
                        }
                    }
                    if (!TAN) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
                    } else if (d == 0) {
                        uint32_t word = 0;
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            word = __funnelshift_l(__float_as_uint(v[e]), word, 1);     // sign of v[e] ends up at bit 31 - e
                            v[e] = fmaxf(v[e], 0.f);
                        }
#pragma unroll
                        for (int jj = 0; jj < MAX_NCH; ++jj) if (jj == j) gate[jj] = word;
                    } else {
                        uint32_t word = 0;
#pragma unroll
                        for (int jj = 0; jj < MAX_NCH; ++jj) if (jj == j) word = gate[jj];
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = (word & (0x80000000u >> e)) ? 0.f : v[e];
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h < nk) {
                            mbar_wait(&a1_empty[t * NS + slot], sph);
                            const uint32_t A1 = smem_u32(smem + C::A1_OFF + (t * NS + slot) * A1_SLOT) + row_off;
                            uint32_t a0[8], a1[8];            // fp16 pairs: a0 = fp16(v), a1 = fp16(v - a0)
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float v0 = v[16 * h + 2 * e], v1 = v[16 * h + 2 * e + 1];
                                a0[e] = pack_f16(v0, v1);
                                const float2 back = unpack_f16(a0[e]);
                                a1[e] = pack_f16(v0 - back.x, v1 - back.y);
                            }
#pragma unroll
                            for (int c = 0; c < 2; ++c) {     // 16-byte chunks 0,1 = a0[0..7], a0[8..15]; 2,3 = a1
                                sts128(A1 + ((c ^ row_sw) << 4), a0[4 * c], a0[4 * c + 1], a0[4 * c + 2], a0[4 * c + 3]);
                                sts128(A1 + (((2 + c) ^ row_sw) << 4), a1[4 * c], a1[4 * c + 1], a1[4 * c + 2], a1[4 * c + 3]);
                            }
                            fence_async_smem();
                            warp_arrive_leader(&a1_full[t * NS + slot], rank, lane);
                            if (++slot == NS) { slot = 0; sph ^= 1; }
                        }
                    }
                    pos += nk;
                }
                if (++d == RPP) { d = 0; ++ks; }
            }
        } else {
            // ---------------- epilogue: layer-1 accumulator -> ReLU -> output layer -> X', dX'/d(X,u) ----------------
            // 16x256b TMEM fragments: this thread holds rows 32w + 16hf + 8hb + lane/4 (hf, hb in {0,1}) and, of
            // every group of 8 columns, columns 2*(lane%4) and +1 -- so one pair of LDS.128 of the output weights
            // feeds 4 rows (a thread-per-row layout needs a broadcast LDS.128 per column: 4x the shared-memory
            // wavefronts, and this kernel is bound by the shared-memory pipe).  Column sums are completed
            // across the 4 lanes of a quad with two shuffles per value.
            const uint32_t W2s = smem_u32(smem + C::W2_OFF + t * C::W2_BYTES);
            const int q4 = lane & 3, rq = lane >> 2;
            int curp = -1;
            uint32_t w2loads = 0;
            uint32_t gate[7];                  // sign bits of this thread's 208 primal pre-activations (TAN)
#pragma unroll
            for (int j = 0; j < 7; ++j) gate[j] = 0;
            int ks = 0, d = 0;
            for (int k = 0; k < nt; ++k) {
                const long long tau = T0 + 2 * ks + t;
                const int p = (int)(tau / tiles_p), l = (int)(tau - (long long)p * tiles_p);
                if (p != curp) {
                    named_bar_sync(1 + t, 128);      // every warp of the team is done with the old weights
                    if (r == 0) {
                        mbar_expect_tx(&w2_full[t], C::W2_BYTES);
                        bulk_g2s(smem + C::W2_OFF + t * C::W2_BYTES, im.W2p + (size_t)p * TILE_N * DP, C::W2_BYTES, &w2_full[t]);
                    }
                    mbar_wait(&w2_full[t], w2loads & 1);
                    ++w2loads;
                    curp = p;
                }
                // D <= 4: packed accumulators (even column, odd column) for FFMA2; wider outputs would spill them
                constexpr bool PK = D <= 4;
                unsigned long long y2[PK ? 4 : 1][PK ? D : 1];
                float y[4][D];
#pragma unroll
                for (int jr = 0; jr < 4; ++jr)
#pragma unroll
                    for (int o = 0; o < D; ++o) {
                        y[jr][o] = 0.f;
                        if constexpr (PK) y2[jr][o] = 0ull;
                    }
                mbar_wait(&acc1_full[t], (uint32_t)k & 1);
                tc_fence_after();
                // half = 16 columns (two 8-column groups) of both row halves; consumes fragment buffer fb and
                // shifts / tests bits [ebase, ebase+16) of the current gate word
                auto half = [&](const int c0, float (&fb)[2][8], uint32_t& word, auto ebase_tag, auto ne_tag) {
                    constexpr int EBASE = decltype(ebase_tag)::value, NE = decltype(ne_tag)::value;
                    if constexpr (PK) {
#pragma unroll
                        for (int gi = 0; gi < 2; ++gi) {
                            // output weights of this thread's column pair, interleaved by prep_w2_kernel as [o][cc]: one 64-bit
                            // operand of FFMA2 per output (packed fp32 FMA: both columns of the pair in one instruction -- the
                            // output layer is 832 FMAs per thread and tile, and the epilogue's duration is on the critical
                            // path of a track: layer-1 MMAs of the next tile wait for it to release the accumulator)
                            unsigned long long w2p[DP];
#pragma unroll
                            for (int o2 = 0; o2 < DP / 2; ++o2) {
                                const float4 q = lds128f(W2s + (uint32_t)((((c0 + 8 * gi) / 2 + q4) * (2 * DP) + 4 * o2) * 4));
                                w2p[2 * o2] = pack2(q.x, q.y);
                                w2p[2 * o2 + 1] = pack2(q.z, q.w);
                            }
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb) {
                                    float vv[2];
#pragma unroll
                                    for (int cc = 0; cc < 2; ++cc) {
                                        vv[cc] = fb[hf][4 * gi + 2 * hb + cc];
                                        const int e = EBASE + (gi * 2 + hf) * 4 + 2 * hb + cc;
                                        if (!TAN) {
                                            vv[cc] = fmaxf(vv[cc], 0.f);
                                        } else if (d == 0) {
                                            word = __funnelshift_l(__float_as_uint(vv[cc]), word, 1);   // element e ends at bit NE-1-e
                                            vv[cc] = fmaxf(vv[cc], 0.f);
                                        } else {
                                            vv[cc] = (word & (1u << (NE - 1 - e))) ? 0.f : vv[cc];
                                        }
                                    }
                                    const unsigned long long vp = pack2(vv[0], vv[1]);
#pragma unroll
                                    for (int o = 0; o < D; ++o) y2[2 * hf + hb][o] = ffma2(vp, w2p[o], y2[2 * hf + hb][o]);
                                }
                        }
                    } else {
#pragma unroll
                        for (int gi = 0; gi < 2; ++gi) {
                            float w2[2][DP];                                        // output weights of this thread's 2 columns
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                                for (int o4 = 0; o4 < DP / 4; ++o4)
                                    *reinterpret_cast<float4*>(&w2[cc][4 * o4]) =
                                        lds128f(W2s + (uint32_t)(((c0 + 8 * gi + 2 * q4 + cc) * DP + 4 * o4) * 4));
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb)
#pragma unroll
                                    for (int cc = 0; cc < 2; ++cc) {
                                        float vv = fb[hf][4 * gi + 2 * hb + cc];
                                        const int e = EBASE + (gi * 2 + hf) * 4 + 2 * hb + cc;
                                        if (!TAN) {
                                            vv = fmaxf(vv, 0.f);
                                        } else if (d == 0) {
                                            word = __funnelshift_l(__float_as_uint(vv), word, 1);   // element e ends at bit NE-1-e
                                            vv = fmaxf(vv, 0.f);
                                        } else {
                                            vv = (word & (1u << (NE - 1 - e))) ? 0.f : vv;
                                        }
#pragma unroll
                                        for (int o = 0; o < D; ++o) y[2 * hf + hb][o] += vv * w2[cc][o];
                                    }
                        }
                    }
                };
                auto issue = [&](const int c0, float (&fb)[2][8]) {
                    const uint32_t ta = track_taddr + ((uint32_t)(w * 32) << 16) + TM_ACC1 + c0;
                    tc_ld_16x256b_x2(ta, fb[0]);
                    tc_ld_16x256b_x2(ta + (16u << 16), fb[1]);
                };
                auto push = [&](uint32_t word) {   // primal pass: push the new word; tangent passes: rotate (oldest = next)
#pragma unroll
                    for (int j = 6; j > 0; --j) gate[j] = gate[j - 1];
                    gate[0] = word;
                };
                // 13 sixteen-column batches, two per trip of a ROLLED loop (a fully unrolled epilogue is 1 400 SASS
                // instructions: with five roles resident, instruction-cache misses were 26 % of its stall samples),
                // software pipelined over two fragment buffers: the TMEM loads of the next batch are issued right
                // after the wait for the current one (tcgen05.wait::ld waits for everything outstanding).
                float fa[2][8], fb2[2][8];
                issue(0, fa);
#pragma unroll 1
                for (int c0 = 0; c0 < (TILE_N / 32) * 32; c0 += 32) {
                    uint32_t word = (TAN && d != 0) ? gate[6] : 0u;
                    tc_wait_ld8(fa[0]); tc_wait_ld8(fa[1]);
                    issue(c0 + 16, fb2);
                    half(c0, fa, word, std::integral_constant<int, 0>(), std::integral_constant<int, 32>());
                    tc_wait_ld8(fb2[0]); tc_wait_ld8(fb2[1]);
                    issue(c0 + 32, fa);
                    half(c0 + 16, fb2, word, std::integral_constant<int, 16>(), std::integral_constant<int, 32>());
                    if (TAN) push(word);
                }
                {
                    static_assert(TILE_N % 32 == 16, "epilogue tail handles one 16-column batch");
                    uint32_t word = (TAN && d != 0) ? gate[6] : 0u;
                    tc_wait_ld8(fa[0]); tc_wait_ld8(fa[1]);
                    half((TILE_N / 32) * 32, fa, word, std::integral_constant<int, 0>(), std::integral_constant<int, 16>());
                    if (TAN) push(word);
                }
                tc_fence_before();
                warp_arrive_leader(&acc1_empty[t], rank, lane);
                // complete the column sums across the quad, then lane q4 writes outputs o = q4 (and q4 + 4)
#pragma unroll
                for (int jr = 0; jr < 4; ++jr)
#pragma unroll
                    for (int o = 0; o < D; ++o) {
                        if constexpr (PK) y[jr][o] = lo2(y2[jr][o]) + hi2(y2[jr][o]);
                        y[jr][o] += __shfl_xor_sync(0xffffffffu, y[jr][o], 1);
                        y[jr][o] += __shfl_xor_sync(0xffffffffu, y[jr][o], 2);
                    }
#pragma unroll
                for (int oo = 0; oo < (D + 3) / 4; ++oo) {
                    const int o = q4 + 4 * oo;
                    if (o < D) {
                        const float sd = n.dX_std ? n.dX_std[o] : 1.f, mn = n.dX_mean ? n.dX_mean[o] : 0.f, bo = n.b2[o];
#pragma unroll
                        for (int jr = 0; jr < 4; ++jr) {
                            const int i = l * (2 * TILE_M) + (int)rank * TILE_M + w * 32 + (jr >> 1) * 16 + (jr & 1) * 8 + rq;
                            if (i < S) {
                                float yo = 0.f;
#pragma unroll
                                for (int o2 = 0; o2 < D; ++o2) if (o2 == o) yo = y[jr][o2];
                                const size_t g = (size_t)i * P + p;
                                if (!TAN || d == 0) a.Xn[g * D + o] = __ldg(a.X + g * D + o) + ((yo + bo) * sd + mn);
                                else a.Jp[(g * D + o) * TD + (d - 1)] = ((d - 1) == o ? 1.f : 0.f) + yo * sd;
                            }
                        }
                    }
                }
                if (++d == RPP) { d = 0; ++ks; }
            }
        }
    }
done:
    __syncwarp();
    tc_fence_before();
    cluster_sync();                            // the leader's MMAs read the peer's shared memory and write its TMEM
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

}  // namespace tc
}  // namespace pddp
