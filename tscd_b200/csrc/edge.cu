// WaveletsHFBlock at the SELECTED anchors only (conv-tower seam).
//
// Reference: the head runs edge_enhance_reg[k] = WaveletsHFBlock(256) densely over every level's regression feature map
// (yolox/models/tscd_head.py:367, yolox/models/surrounding_extraction.py:215-267) and then keeps only the rows of the selected
// anchors (find_feature_score, tscd_head.py:976-991): for 576 x 576 inputs that is a 3x3 conv (256->256) and a 1x1 conv
// (768->768) over 6804 positions per frame to use 30..500 of them.  Every output position depends on a 3x3 neighbourhood only:
//   x_content(y,x) = ReLU(W3 * patch3x3(y,x) + b3)                                  (filter2)
//   HF(Y,X)        = Haar high-pass sub-bands [LH|HL|HH] of the 2x2 block (Y,X) = (y/2, x/2)   (DWT_2D, stride-2 conv)
//   HF'(Y,X)       = ReLU(W1 * HF(Y,X) + b1)                                        (filter1)
//   x_idwt(y,x)    = 1/2 (s_lh LH' + s_hl HL' + s_hh HH'),  signs from (y&1, x&1), LL = 0      (IDWT_2D, stride-2 transposed conv)
//   edge(y,x)      = x_content(y,x) * x_idwt(y,x)
// so the block is evaluated per bank row:
//   edge_patch_kernel    writes, for every kept proposal, its 3x3 patch [9 x 256] and its Haar sub-bands [3 x 256] as 16-bit rows,
//                        grouped by pyramid level (each level has its own conv weights) into per-level segments;
//   tscd_linear          (tcgen05 GEMM, host-issued per level with the device-side row count) applies W3 / W1 + bias;
//   edge_combine_kernel  ReLUs, inverse transform signs, product -> the bank's edge rows.
#include "common.cuh"

namespace tscd {

constexpr int kEdgeThreads = 256;
constexpr int kEdgeChunk = 512;              // rows of a frame whose level ranks are resolved per pass
constexpr int kEdgeDim = 256;

// 8 consecutive channels (lane * 8 ..) of one pixel as fp32; p = first channel of the pixel
template <typename TF>
__device__ __forceinline__ void load_chan8(const TF* p, int64_t cs, int lane, float (&v)[8]) {
    if (cs == 1 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        if constexpr (sizeof(TF) == 2) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p + lane * 8));
            const TF* e = reinterpret_cast<const TF*>(&raw);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ldf_reg(e[k]);
        } else {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p + lane * 8));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p + lane * 8) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ldf(p + (int64_t)(lane * 8 + k) * cs);
    }
}

template <typename TB>
__device__ __forceinline__ void store_chan8(TB* dst, int lane, const float (&v)[8]) {
    uint4 out;
    TB* o = reinterpret_cast<TB*>(&out);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = cvt_from_float<TB>(v[k]);
    *reinterpret_cast<uint4*>(dst + lane * 8) = out;
}

template <typename TF, typename TB>
__global__ void __launch_bounds__(kEdgeThreads) edge_patch_kernel(const tscd_edge_patches_args a) {
    __shared__ int s_cnt[TSCD_MAX_LEVELS], s_base[TSCD_MAX_LEVELS];
    __shared__ unsigned short s_rank[kEdgeChunk];
    const int frame = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = kEdgeThreads >> 5;
    const int n = min(a.sel_count[frame], a.max_keep);
    const int row0 = a.row_off[frame];
    const int32_t* idx = a.sel_idx + (int64_t)frame * a.max_keep;
    TB* patches = reinterpret_cast<TB*>(a.patches);
    TB* hf = reinterpret_cast<TB*>(a.hf);
    for (int c0 = 0; c0 < n; c0 += kEdgeChunk) {
        const int m = min(kEdgeChunk, n - c0);
        __syncthreads();
        if (tid < TSCD_MAX_LEVELS) s_cnt[tid] = 0;
        __syncthreads();
        for (int jj = tid; jj < m; jj += kEdgeThreads) {
            const AnchorPos p = anchor_pos(a.anchors, idx[c0 + jj]);
            s_rank[jj] = (unsigned short)atomicAdd(&s_cnt[p.level], 1);
        }
        __syncthreads();
        // one global atomic per level and chunk: where this chunk's rows start inside the level's segment
        if (tid < a.anchors.num_levels) s_base[tid] = s_cnt[tid] ? atomicAdd(&a.level_count[tid], s_cnt[tid]) : 0;
        __syncthreads();
        for (int jj = wid; jj < m; jj += nw) {
            const AnchorPos p = anchor_pos(a.anchors, idx[c0 + jj]);
            const int l = p.level, W = a.anchors.level_w[l], H = a.anchors.level_h[l];
            const int y = p.local / W, x = p.local - y * W;
            const int rank = s_base[l] + (int)s_rank[jj];
            if (rank >= a.seg_cap[l]) {
                if (lane == 0) {
                    if (a.status) atomicMin(a.status, TSCD_ERR_CAPACITY);
                    a.slot[row0 + c0 + jj] = -1;
                }
                continue;
            }
            const int64_t slot = a.seg_base[l] + rank;
            if (lane == 0) a.slot[row0 + c0 + jj] = (int)(slot * 4 + (y & 1) * 2 + (x & 1));
            const TF* lvl = reinterpret_cast<const TF*>(a.feat_reg.ptr[l]) + (int64_t)frame * a.feat_reg.frame_stride[l];
            const int64_t as = a.feat_reg.anchor_stride[l], cs = a.feat_reg.chan_stride[l];
            // ---- 3x3 patch, zero padded (filter2: Conv2d(256, 256, 3, padding=1)), tap-major ----
            TB* prow = patches + slot * (9 * kEdgeDim);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
                float v[8];
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) load_chan8<TF>(lvl + (int64_t)(yy * W + xx) * as, cs, lane, v);
                else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = 0.f;
                }
                store_chan8<TB>(prow + t * kEdgeDim, lane, v);
            }
            // ---- Haar high-pass sub-bands of the anchor's 2x2 block (DWT_2D: w[i][j], stride 2, cross-correlation) ----
            const int y0 = y & ~1, x0 = x & ~1;
            float p00[8], p01[8], p10[8], p11[8], lh[8], hl[8], hh[8];
            load_chan8<TF>(lvl + (int64_t)(y0 * W + x0) * as, cs, lane, p00);
            load_chan8<TF>(lvl + (int64_t)(y0 * W + x0 + 1) * as, cs, lane, p01);
            load_chan8<TF>(lvl + (int64_t)((y0 + 1) * W + x0) * as, cs, lane, p10);
            load_chan8<TF>(lvl + (int64_t)((y0 + 1) * W + x0 + 1) * as, cs, lane, p11);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                lh[k] = 0.5f * (p00[k] + p01[k] - p10[k] - p11[k]);    // low along x, high along y
                hl[k] = 0.5f * (p00[k] - p01[k] + p10[k] - p11[k]);    // high along x, low along y
                hh[k] = 0.5f * (p00[k] - p01[k] - p10[k] + p11[k]);
            }
            TB* hrow = hf + slot * (3 * kEdgeDim);
            store_chan8<TB>(hrow, lane, lh);
            store_chan8<TB>(hrow + kEdgeDim, lane, hl);
            store_chan8<TB>(hrow + 2 * kEdgeDim, lane, hh);
        }
    }
}

// edge = ReLU(content) * round(1/2 (s_lh ReLU(LH') + s_hl ReLU(HL') + s_hh ReLU(HH')))   -- one warp per bank row
template <typename TB>
__global__ void __launch_bounds__(kEdgeThreads) edge_combine_kernel(const tscd_edge_combine_args a) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (kEdgeThreads >> 5) + (threadIdx.x >> 5);
    if (r >= *a.total_rows || r >= a.rows_cap) return;
    const int sp = a.slot[r];
    if (sp < 0) return;
    const int64_t slot = sp >> 2;
    const float s_lh = (sp & 2) ? -1.f : 1.f, s_hl = (sp & 1) ? -1.f : 1.f, s_hh = s_lh * s_hl;
    const TB* c = reinterpret_cast<const TB*>(a.content) + slot * kEdgeDim;
    const TB* h = reinterpret_cast<const TB*>(a.hf_out) + slot * (3 * kEdgeDim);
    float vc[8], v0[8], v1[8], v2[8], o[8];
    load_chan8<TB>(c, 1, lane, vc);
    load_chan8<TB>(h, 1, lane, v0);
    load_chan8<TB>(h + kEdgeDim, 1, lane, v1);
    load_chan8<TB>(h + 2 * kEdgeDim, 1, lane, v2);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float idwt = 0.5f * (s_lh * fmaxf(v0[k], 0.f) + s_hl * fmaxf(v1[k], 0.f) + s_hh * fmaxf(v2[k], 0.f));
        // the reference holds x_idwt in the evaluation dtype before the product
        o[k] = fmaxf(vc[k], 0.f) * ldf_reg(cvt_from_float<TB>(idwt));
    }
    store_chan8<TB>(reinterpret_cast<TB*>(a.bank_edge) + (int64_t)r * kEdgeDim, lane, o);
}

template <typename TF>
static int launch_patch(const tscd_edge_patches_args* a, cudaStream_t st) {
    if (a->op_dtype == TSCD_F16) edge_patch_kernel<TF, __half><<<a->num_frames, kEdgeThreads, 0, st>>>(*a);
    else edge_patch_kernel<TF, __nv_bfloat16><<<a->num_frames, kEdgeThreads, 0, st>>>(*a);
    return TSCD_OK;
}

}  // namespace tscd

extern "C" int tscd_edge_patches(const tscd_edge_patches_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->num_frames <= 0 || a->max_keep <= 0 || !a->sel_idx || !a->sel_count || !a->row_off || !a->level_count || !a->slot ||
        !a->patches || !a->hf)
        return TSCD_ERR_INVALID_ARG;
    if (a->op_dtype != TSCD_F16 && a->op_dtype != TSCD_BF16) return TSCD_ERR_UNSUPPORTED;
    if (a->anchors.num_levels <= 0 || a->anchors.num_levels > TSCD_MAX_LEVELS) return TSCD_ERR_INVALID_ARG;
    for (int l = 0; l < a->anchors.num_levels; ++l) {
        // the reference's stride-2 transform pairs rows / columns: an odd map has no inverse of the same size (it raises there too)
        if ((a->anchors.level_h[l] & 1) || (a->anchors.level_w[l] & 1)) return TSCD_ERR_UNSUPPORTED;
        if (!a->feat_reg.ptr[l] || a->seg_cap[l] <= 0 || a->seg_base[l] < 0) return TSCD_ERR_INVALID_ARG;
        if ((int64_t)a->seg_base[l] + a->seg_cap[l] > (1 << 28)) return TSCD_ERR_CAPACITY;
    }
    if ((reinterpret_cast<uintptr_t>(a->patches) | reinterpret_cast<uintptr_t>(a->hf)) & 15) return TSCD_ERR_INVALID_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(a->level_count, 0, sizeof(int32_t) * TSCD_MAX_LEVELS, st) != cudaSuccess) return TSCD_ERR_CUDA;
    int rc;
    switch (a->feat_dtype) {
        case TSCD_F32: rc = launch_patch<float>(a, st); break;
        case TSCD_F16: rc = launch_patch<__half>(a, st); break;
        case TSCD_BF16: rc = launch_patch<__nv_bfloat16>(a, st); break;
        default: return TSCD_ERR_UNSUPPORTED;
    }
    if (rc != TSCD_OK) return rc;
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}

extern "C" int tscd_edge_combine(const tscd_edge_combine_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->rows_cap <= 0 || !a->total_rows || !a->slot || !a->content || !a->hf_out || !a->bank_edge) return TSCD_ERR_INVALID_ARG;
    if (a->op_dtype != TSCD_F16 && a->op_dtype != TSCD_BF16) return TSCD_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(a->content) | reinterpret_cast<uintptr_t>(a->hf_out) | reinterpret_cast<uintptr_t>(a->bank_edge)) & 15)
        return TSCD_ERR_INVALID_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (a->rows_cap + (kEdgeThreads >> 5) - 1) / (kEdgeThreads >> 5);
    if (a->op_dtype == TSCD_F16) edge_combine_kernel<__half><<<grid, kEdgeThreads, 0, st>>>(*a);
    else edge_combine_kernel<__nv_bfloat16><<<grid, kEdgeThreads, 0, st>>>(*a);
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}
