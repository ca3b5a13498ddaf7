// Cosine multi-head attention over ragged per-item row sets on warp-level tensor cores (mma.sync m16n8k16, fp32 accumulate):
// the per-frame attentions of the stage when a frame holds MORE than 32 proposals (mode B: 50..500 per frame).
//
// Reference: PositionMHAttention.forward (yolox/models/tscd_matching.py:31-60; CAFM, 8 heads x 32) and MHAttention.forward
// (:159-181; TaskAligned, 8 heads x 128): q, k L2-normalised per head, NO scale, softmax(q k^T) v.  Queries and keys / values of
// item i are the row ranges [q_beg[i], q_end[i]) of q and [kv_beg[i], kv_end[i]) of k / v (read from device memory: no host
// sync), so the same kernel serves TaskAligned (queries = keys = the rows of local frame i) and the wide CAFM chain step
// (queries = a clip's matched previous-frame rows, keys / values = the current frame's rows).
//
// One CTA = (64 queries, head, item), 4 warps x 16 query rows, flash-style loop over 64-key tiles: S = Q K^T on the raw 16-bit
// operands, scaled by 1/|q_r| * 1/|k_j| in the fp32 accumulators; the logits are cosines in [-1, 1], so exp() needs no running
// maximum and no rescaling of the output accumulator; P (16-bit) feeds P @ V straight from the accumulator registers.
// Frames of <= 32 rows keep their single-tile kernels (csrc/tail.cu frame_attention16_kernel, csrc/cafm.cu fast chain).
#include "common.cuh"
#include "mma.cuh"

namespace tscd {

constexpr int kFlashBQ = 64, kFlashBK = 64, kFlashThreads = 128;

template <typename T, int HD>
__global__ void __launch_bounds__(kFlashThreads) frame_flash_kernel(const tscd_frame_flash_args a) {
    constexpr int P = HD + 8;                       // padded row pitch (16-bit elements): ldmatrix conflict-free
    constexpr int CH = HD / 8;                      // 16-byte chunks per row
    extern __shared__ __align__(16) unsigned char ff_smem[];
    T* sQ = reinterpret_cast<T*>(ff_smem);          // [64][P]
    T* sK = sQ + kFlashBQ * P;
    T* sV = sK + kFlashBK * P;
    float* sQn = reinterpret_cast<float*>(sV + kFlashBK * P);   // [64] 1/|q|
    float* sKn = sQn + kFlashBQ;                                // [64] 1/|k|
    const int item = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kFlashBQ;
    const int qb = a.q_beg[item], nq = a.q_end[item] - qb;
    const int kb = a.kv_beg[item], nk = a.kv_end[item] - kb;
    if (q0 >= nq || nk <= 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const T* q = reinterpret_cast<const T*>(a.q) + h * HD;
    const T* k = reinterpret_cast<const T*>(a.k) + h * HD;
    const T* v = reinterpret_cast<const T*>(a.v) + h * HD;

    for (int i = tid; i < kFlashBQ * CH; i += kFlashThreads) {
        const int r = i / CH, c = i - r * CH;
        if (q0 + r < nq) cp_async16(sQ + r * P + c * 8, q + (int64_t)(qb + q0 + r) * a.ldq + c * 8);
        else *reinterpret_cast<uint4*>(sQ + r * P + c * 8) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
    float o[HD / 8][4];
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
    float l0 = 0.f, l1 = 0.f;
    float iq0 = 0.f, iq1 = 0.f;

    for (int k0 = 0; k0 < nk; k0 += kFlashBK) {
        __syncthreads();                            // previous tile fully consumed
        for (int i = tid; i < kFlashBK * CH; i += kFlashThreads) {
            const int r = i / CH, c = i - r * CH;
            if (k0 + r < nk) {
                cp_async16(sK + r * P + c * 8, k + (int64_t)(kb + k0 + r) * a.ldk + c * 8);
                cp_async16(sV + r * P + c * 8, v + (int64_t)(kb + k0 + r) * a.ldv + c * 8);
            } else {                                // padding keys: masked below; values must be finite (x 0)
                *reinterpret_cast<uint4*>(sK + r * P + c * 8) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(sV + r * P + c * 8) = make_uint4(0, 0, 0, 0);
            }
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        // inverse L2 norms per head: thread r < 64 -> key r of the tile (and, once, query r)
        if (tid < kFlashBK) {
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                float x[8];
                load8(sK + tid * P + c * 8, x);
#pragma unroll
                for (int i = 0; i < 8; ++i) ss = fmaf(x[i], x[i], ss);
            }
            sKn[tid] = 1.f / sqrtf(ss);
        } else if (k0 == 0) {
            const int r = tid - kFlashBK;
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                float x[8];
                load8(sQ + r * P + c * 8, x);
#pragma unroll
                for (int i = 0; i < 8; ++i) ss = fmaf(x[i], x[i], ss);
            }
            sQn[r] = 1.f / sqrtf(ss);
        }
        __syncthreads();
        if (k0 == 0) { iq0 = sQn[warp * 16 + g]; iq1 = sQn[warp * 16 + g + 8]; }

        // ---- S = Q K^T (16 query rows x 64 keys per warp) ----
        float sc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t qa[4];
            ldsm_x4(qa, sQ + (warp * 16 + (lane & 15)) * P + ks * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int ntp = 0; ntp < 4; ++ntp) {     // two key n-tiles per ldmatrix.x4
                uint32_t kf[4];
                ldsm_x4(kf, sK + (ntp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * P + ks * 16 + ((lane >> 3) & 1) * 8);
                mma16816<T>(sc[2 * ntp], qa, kf[0], kf[1]);
                mma16816<T>(sc[2 * ntp + 1], qa, kf[2], kf[3]);
            }
        }
        // ---- P = exp(cos), masked; row sums ----
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * t4;
            const float ik0 = sKn[c], ik1 = sKn[c + 1];
            const bool v0 = k0 + c < nk, v1 = k0 + c + 1 < nk;
            sc[nt][0] = v0 ? __expf(sc[nt][0] * iq0 * ik0) : 0.f;
            sc[nt][1] = v1 ? __expf(sc[nt][1] * iq0 * ik1) : 0.f;
            sc[nt][2] = v0 ? __expf(sc[nt][2] * iq1 * ik0) : 0.f;
            sc[nt][3] = v1 ? __expf(sc[nt][3] * iq1 * ik1) : 0.f;
            l0 += sc[nt][0] + sc[nt][1];
            l1 += sc[nt][2] + sc[nt][3];
        }
        // ---- O += P V ----
#pragma unroll
        for (int j = 0; j < 4; ++j) {               // 16 keys per step
            uint32_t pa[4];
            pa[0] = pack2<T>(sc[2 * j][0], sc[2 * j][1]);
            pa[1] = pack2<T>(sc[2 * j][2], sc[2 * j][3]);
            pa[2] = pack2<T>(sc[2 * j + 1][0], sc[2 * j + 1][1]);
            pa[3] = pack2<T>(sc[2 * j + 1][2], sc[2 * j + 1][3]);
#pragma unroll
            for (int dp = 0; dp < HD / 16; ++dp) {  // two 8-wide output tiles per ldmatrix.x4.trans
                uint32_t vf[4];
                ldsm_x4_trans(vf, sV + (j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * P + dp * 16 + ((lane >> 4) & 1) * 8);
                mma16816<T>(o[2 * dp], pa, vf[0], vf[1]);
                mma16816<T>(o[2 * dp + 1], pa, vf[2], vf[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float r0 = 1.f / l0, r1 = 1.f / l1;
    const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
        const int c = h * HD + d * 8 + 2 * t4;
        if (row0 < nq) *reinterpret_cast<float2*>(a.out + (int64_t)(qb + row0) * a.ldo + c) = make_float2(o[d][0] * r0, o[d][1] * r0);
        if (row1 < nq) *reinterpret_cast<float2*>(a.out + (int64_t)(qb + row1) * a.ldo + c) = make_float2(o[d][2] * r1, o[d][3] * r1);
    }
}

template <typename T, int HD>
static int launch_flash(const tscd_frame_flash_args* a, cudaStream_t st) {
    const size_t sm = (size_t)(kFlashBQ + 2 * kFlashBK) * (HD + 8) * 2 + (kFlashBQ + kFlashBK) * sizeof(float);
    if (cudaFuncSetAttribute(frame_flash_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return TSCD_ERR_CUDA;
    const dim3 grid((a->max_q + kFlashBQ - 1) / kFlashBQ, a->heads, a->num_items);
    frame_flash_kernel<T, HD><<<grid, kFlashThreads, sm, st>>>(*a);
    return TSCD_OK;
}

}  // namespace tscd

extern "C" int tscd_frame_flash(const tscd_frame_flash_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->num_items <= 0 || a->heads <= 0 || a->max_q <= 0 || !a->q_beg || !a->q_end || !a->kv_beg || !a->kv_end || !a->q || !a->k ||
        !a->v || !a->out)
        return TSCD_ERR_INVALID_ARG;
    if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->ldo % 2)) return TSCD_ERR_INVALID_ARG;
    if (a->num_items > 65535 || a->heads > 65535) return TSCD_ERR_CAPACITY;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    if (a->dtype == TSCD_F16) {
        if (a->head_dim == 32) rc = launch_flash<__half, 32>(a, st);
        else if (a->head_dim == 128) rc = launch_flash<__half, 128>(a, st);
        else return TSCD_ERR_UNSUPPORTED;
    } else if (a->dtype == TSCD_BF16) {
        if (a->head_dim == 32) rc = launch_flash<__nv_bfloat16, 32>(a, st);
        else if (a->head_dim == 128) rc = launch_flash<__nv_bfloat16, 128>(a, st);
        else return TSCD_ERR_UNSUPPORTED;
    } else {
        return TSCD_ERR_UNSUPPORTED;
    }
    if (rc != TSCD_OK) return rc;
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}
