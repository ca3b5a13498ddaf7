// TMEM read probe: how much does a tcgen05.ld cost as a function of its width, and how do the loads of several warps share
// the read port?  One CTA per SM allocates 512 columns; W warps (warp w reads lane quadrant w % 4) each issue `iters` loads of
// `cols` consecutive columns (x16 / x32 / x64 / x128) followed by tcgen05.wait::ld; cycles per CTA / bytes per clock are printed.
// Decides how the softmax warps of the attention kernels should split a score tile (16 warps x 16 columns vs 4 warps x 64).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_probe tools/tmem_probe.cu && tools/bin/tmem_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t tmem_ld_sum(uint32_t taddr);

#define LD_BODY(X, NREG)                                                                                                   \
    template <> __device__ __forceinline__ uint32_t tmem_ld_sum<X>(uint32_t taddr) {                                       \
        uint32_t r[NREG];                                                                                                   \
        LD_ASM_##X                                                                                                          \
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");                                                     \
        uint32_t s = 0;                                                                                                     \
        _Pragma("unroll") for (int i = 0; i < NREG; ++i) s ^= r[i];                                                         \
        return s;                                                                                                           \
    }

#define R4(b) "=r"(r[b]), "=r"(r[b + 1]), "=r"(r[b + 2]), "=r"(r[b + 3])
#define R16(b) R4(b), R4(b + 4), R4(b + 8), R4(b + 12)
#define LD_ASM_16                                                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n" \
                 : R16(0) : "r"(taddr) : "memory");
#define LD_ASM_32                                                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n" \
                 : R16(0), R16(16) : "r"(taddr) : "memory");
#define LD_ASM_64                                                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n" \
                 : R16(0), R16(16), R16(32), R16(48) : "r"(taddr) : "memory");
LD_BODY(16, 16)
LD_BODY(32, 32)
LD_BODY(64, 64)

template <int X>
__global__ void __launch_bounds__(512, 1) probe(int warps, int iters, unsigned long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    uint32_t acc = 0;
    const long long t0 = clock64();
    if (warp < warps) {
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int group = warp >> 2;                       // warps of one lane quadrant read different column ranges
        for (int i = 0; i < iters; ++i) {
            const uint32_t col = (uint32_t)(((group * X) + i * X * ((warps + 3) / 4)) & (512 - X)) & ~(uint32_t)(X - 1);
            acc ^= tmem_ld_sum<X>(lane_base + col);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
    }
}

template <int X>
static void run(int warps, int iters, unsigned long long* d_cyc, uint32_t* d_sink, int ctas) {
    probe<X><<<ctas, 512>>>(warps, iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    probe<X><<<ctas, 512>>>(warps, iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    unsigned long long h[256];
    CK(cudaMemcpy(h, d_cyc, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < ctas; ++i) avg += (double)h[i];
    avg /= ctas;
    const double bytes = (double)warps * iters * 32.0 * X * 4.0;
    printf("{\"cols_per_load\": %d, \"warps\": %d, \"loads_per_warp\": %d, \"cycles\": %.0f, \"cycles_per_load_per_warp\": %.1f, \"bytes_per_clk_per_sm\": %.1f}\n",
           X, warps, iters, avg, avg / iters, bytes / avg);
}

int main() {
    unsigned long long* d_cyc;
    uint32_t* d_sink;
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * 256));
    CK(cudaMalloc(&d_sink, sizeof(uint32_t) * 512));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int ctas = p.multiProcessorCount;
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16}) {
        run<16>(warps, iters, d_cyc, d_sink, ctas);
        run<32>(warps, iters / 2, d_cyc, d_sink, ctas);
        run<64>(warps, iters / 4, d_cyc, d_sink, ctas);
    }
    return 0;
}
