// Probe: which (row, column) of a TMEM tile does each register of tcgen05.ld.16x256b.x8 hold?
// Writes value = row*1000 + col with 32x32b stores (thread = row), reads back with 16x256b.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(float* out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
    // store: thread = row (tid), 64 columns, 16 at a time
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        for (int e = 0; e < 16; ++e) v[e] = __float_as_uint((float)(tid * 1000 + c0 + e));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(taddr + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                     "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + ((uint32_t)(half * 16) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 32; ++e) out[((warp * 2 + half) * 32 + lane) * 32 + e] = __uint_as_float(r[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base));
}
int main() {
    float* d; cudaMalloc(&d, 4 * 2 * 32 * 32 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    static float h[4 * 2 * 32 * 32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int w = 0; w < 2; ++w) for (int half = 0; half < 2; ++half) for (int lane = 0; lane < 32; lane += 1) {
        if (lane > 5 && lane != 31) continue;
        printf("warp %d half %d lane %2d:", w, half, lane);
        for (int e = 0; e < 12; ++e) { int v = (int)h[((w * 2 + half) * 32 + lane) * 32 + e]; printf(" (%d,%d)", v / 1000, v % 1000); }
        printf(" ... r31=(%d,%d)\n", (int)h[((w * 2 + half) * 32 + lane) * 32 + 31] / 1000, (int)h[((w * 2 + half) * 32 + lane) * 32 + 31] % 1000);
    }
    return 0;
}
