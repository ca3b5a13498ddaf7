// Zero-copy probe: how fast can B200 kernels read scattered rows straight out of PINNED HOST memory (UVA, over PCIe)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/zc_probe tools/zc_probe.cu && gpurun_out/zc_probe
//
// Decides the layout of the e2e (host-buffer) path of the stage: forward_host copies the objectness plane with the copy
// engine and lets K1 / K3 fetch only the survivors' rows in place.  For every row size (64 B fused logit rows, 128 B,
// 512 B feature rows) and launch shape it prints rows/s and GB/s, alone and with a concurrent cudaMemcpyAsync H2D stream.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// one thread loads 16 B; a row of RB bytes is read by RB/16 consecutive threads; U independent rows in flight per thread
template <int U>
__global__ void gather_rows(const uint4* __restrict__ src, const uint32_t* __restrict__ idx, uint4* __restrict__ dst,
                            int n_rows, int vec_per_row) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)n_rows * vec_per_row;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = t; i0 < total; i0 += stride * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < total) {
                const int r = (int)(i / vec_per_row), c = (int)(i % vec_per_row);
                v[u] = __ldg(src + (int64_t)idx[r] * vec_per_row + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < total) dst[i] = v[u];
        }
    }
}

// cp.async variants (what select_rows_kernel uses): mode 0 = .cg, RPV adjacent lanes per row; 1 = .ca, adjacent lanes;
// 2 = .cg, one thread copies the 4 vectors of its row with 4 instructions; 3 = LDG.128, one thread per row, 4 loads
__global__ void gather_rows_async(const uint4* __restrict__ src, const uint32_t* __restrict__ idx, uint4* __restrict__ dst,
                                  int n_rows, int mode) {
    __shared__ uint4 stage[512 * 4];
    const int rows_per_cta = (mode >= 2) ? 512 : 128;
    for (int r0 = blockIdx.x * rows_per_cta; r0 < n_rows; r0 += gridDim.x * rows_per_cta) {
        if (mode < 2) {
            const int r = r0 + threadIdx.x / 4, c = threadIdx.x % 4;
            if (r < n_rows) {
                const uint32_t d = (uint32_t)__cvta_generic_to_shared(stage + threadIdx.x);
                const uint4* g = src + (int64_t)idx[r] * 4 + c;
                if (mode == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            if (r < n_rows) dst[(int64_t)r * 4 + c] = stage[threadIdx.x];
        } else {
            const int r = r0 + threadIdx.x;
            if (r < n_rows) {
                const uint4* g = src + (int64_t)idx[r] * 4;
                if (mode == 2) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t d = (uint32_t)__cvta_generic_to_shared(stage + threadIdx.x * 4 + c);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g + c) : "memory");
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[(int64_t)r * 4 + c] = stage[threadIdx.x * 4 + c];
                } else {
                    uint4 v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[c] = __ldg(g + c);
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[(int64_t)r * 4 + c] = v[c];
                }
            }
        }
    }
}

int main() {
    const size_t host_bytes = (size_t)2 << 30;
    uint4* h;
    CK(cudaHostAlloc(&h, host_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    for (size_t i = 0; i < host_bytes / 16; i += 64) h[i] = make_uint4((uint32_t)i, 1, 2, 3);
    uint4 *dst, *dcopy;
    CK(cudaMalloc(&dst, (size_t)256 << 20));
    CK(cudaMalloc(&dcopy, (size_t)1 << 30));
    cudaStream_t s1, s2;
    CK(cudaStreamCreate(&s1));
    CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int row_bytes[3] = {64, 128, 512};
    for (int pattern = 0; pattern < 2; ++pattern)
    for (int rb = 0; rb < 3; ++rb) {
        const int RB = row_bytes[rb], vpr = RB / 16;
        const int n_rows = (int)(((size_t)96 << 20) / RB);           // 96 MB per launch
        const size_t host_rows = host_bytes / RB;
        uint32_t* hidx = (uint32_t*)malloc((size_t)n_rows * 4);
        uint64_t s = 88172645463325252ull;
        for (int i = 0; i < n_rows; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; hidx[i] = (uint32_t)(s % host_rows); }
        if (pattern == 1) {
            // frame-like locality: consecutive groups of 750 rows come from one contiguous block of 6804 rows (one frame's
            // anchors), the way K1's survivors do
            const size_t blocks = host_rows / 6804;
            for (int i = 0; i < n_rows; ++i) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                hidx[i] = (uint32_t)(((size_t)(i / 750) % blocks) * 6804 + s % 6804);
            }
        }
        uint32_t* didx;
        CK(cudaMalloc(&didx, (size_t)n_rows * 4));
        CK(cudaMemcpy(didx, hidx, (size_t)n_rows * 4, cudaMemcpyHostToDevice));
        for (int with_copy = 0; with_copy < 2; ++with_copy) {
            for (int shape = 1; shape < 4; shape += 2) {
                const int ctas = shape == 0 ? 148 : shape == 1 ? 148 * 4 : shape == 2 ? 148 * 8 : 148 * 16;
                const int threads = shape == 0 ? 256 : 512;
                for (int U = 1; U <= 4; U *= 4) {
                    float best = 1e30f;
                    for (int rep = 0; rep < 3; ++rep) {
                        if (with_copy) CK(cudaMemcpyAsync(dcopy, (char*)h + ((size_t)1 << 30), (size_t)1 << 30, cudaMemcpyHostToDevice, s2));
                        CK(cudaEventRecord(e0, s1));
                        if (U == 1) gather_rows<1><<<ctas, threads, 0, s1>>>(h, didx, dst, n_rows, vpr);
                        else gather_rows<4><<<ctas, threads, 0, s1>>>(h, didx, dst, n_rows, vpr);
                        CK(cudaEventRecord(e1, s1));
                        CK(cudaDeviceSynchronize());
                        float ms;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (ms < best) best = ms;
                    }
                    printf("{\"pattern\": \"%s\", \"row_bytes\": %d, \"concurrent_h2d\": %d, \"ctas\": %d, \"threads\": %d, \"loads_in_flight_per_thread\": %d, "
                           "\"ms\": %.3f, \"Mrows_per_s\": %.1f, \"GBps\": %.2f}\n", pattern ? "750 of 6804 consecutive rows" : "uniform over 2 GiB", RB, with_copy, ctas, threads, U, best,
                           n_rows / best / 1e3, (double)n_rows * RB / best / 1e6);
                }
            }
        }
        if (RB == 64 && pattern == 1) {
            const char* names[4] = {"cp.async.cg, 4 adjacent lanes per row", "cp.async.ca, 4 adjacent lanes per row",
                                    "cp.async.cg, one thread issues the 4 vectors of a row", "ld.global.nc.v4 x4, one thread per row"};
            for (int mode = 0; mode < 4; ++mode) {
                float best = 1e30f;
                for (int rep = 0; rep < 3; ++rep) {
                    CK(cudaEventRecord(e0, s1));
                    gather_rows_async<<<148 * 4, 512, 0, s1>>>(h, didx, dst, n_rows, mode);
                    CK(cudaEventRecord(e1, s1));
                    CK(cudaDeviceSynchronize());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                printf("{\"pattern\": \"750 of 6804 consecutive rows\", \"row_bytes\": 64, \"variant\": \"%s\", \"ms\": %.3f, \"Mrows_per_s\": %.1f, \"GBps\": %.2f}\n",
                       names[mode], best, n_rows / best / 1e3, (double)n_rows * RB / best / 1e6);
            }
        }
        CK(cudaFree(didx));
        free(hidx);
    }
    // reference points: plain H2D copy of 1 GiB, alone
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0, s2));
        CK(cudaMemcpyAsync(dcopy, h, (size_t)1 << 30, cudaMemcpyHostToDevice, s2));
        CK(cudaEventRecord(e1, s2));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"h2d_copy_GiB\": 1, \"ms\": %.2f, \"GBps\": %.2f}\n", ms, 1073.741824 / ms);
    }
    return 0;
}
