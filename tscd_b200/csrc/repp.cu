// REPP detection linking on the device: the O(n1 * n2) part of the reference's tubelet post-processing.
//
// Reference: tools/REPP.py:82-133 (get_video_pairs + solve_distances_def) with the linking scores of :52-79
// (distance_def / distance_logreg) and the pair features of tools/repp_utils.py:34-109.  The reference fills an n1 x n2
// distance matrix with a Python double loop (one sklearn predict_proba call per pair) and then repeatedly takes the global
// minimum (first in row-major order among equals), erasing its row and column, until only INF is left.
//
// Here one CTA handles one pair of consecutive frames: every thread evaluates linking distances (same operation order and
// precision as the numpy expressions: box / ratio / IoU arithmetic in fp32, the centre distance's square root, the
// logistic model and the score algebra in fp64), the finite ones are appended to a candidate list in the workspace, and the
// greedy matching is a loop of block-wide arg-min reductions over the surviving candidates with (distance, row-major index)
// as the key -- exactly the order the reference extracts pairs in, which the tubelet builder depends on.
#include "common.cuh"

namespace tscd {

constexpr int kReppThreads = 256;
constexpr int kReppMaxDet = 4096;          // detections per frame (row / column "used" bitmaps live in shared memory)
constexpr double kReppInf = 9e15;          // REPP.py:16

// repp_utils.py:53-109 on (x, y, w, h) fp32 boxes; every operation rounded to fp32 like the numpy float32 scalars
__device__ __forceinline__ float repp_iou(float4 p, float4 q) {
    const float px2 = __fadd_rn(p.z, p.x), py2 = __fadd_rn(p.w, p.y), qx2 = __fadd_rn(q.z, q.x), qy2 = __fadd_rn(q.w, q.y);
    const float xl = fmaxf(p.x, q.x), yt = fmaxf(p.y, q.y), xr = fminf(px2, qx2), yb = fminf(py2, qy2);
    if (xr < xl || yb < yt) return 0.f;
    const float inter = __fmul_rn(__fsub_rn(xr, xl), __fsub_rn(yb, yt));
    const float a1 = __fmul_rn(__fsub_rn(px2, p.x), __fsub_rn(py2, p.y)), a2 = __fmul_rn(__fsub_rn(qx2, q.x), __fsub_rn(qy2, q.y));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
}

__device__ __forceinline__ double repp_distance(const tscd_repp_link_args& a, int i, int j) {
    const float4 p = reinterpret_cast<const float4*>(a.bbox)[i], q = reinterpret_cast<const float4*>(a.bbox)[j];
    const double si = a.score[i], sj = a.score[j];
    const double dot = a.cls[i] == a.cls[j] ? __dmul_rn(si, sj) : 0.0;       // one-hot score vectors (REPP.py:251-253)
    const float iou = repp_iou(p, q);
    if (a.distance_func == 0) {                                               // distance_def :52-57
        const double div = __dmul_rn((double)iou, dot);
        return div == 0.0 ? kReppInf : __ddiv_rn(1.0, div);
    }
    const float wrel = __fdiv_rn(fminf(p.z, q.z), fmaxf(p.z, q.z)), hrel = __fdiv_rn(fminf(p.w, q.w), fmaxf(p.w, q.w));
    const float2 c1 = reinterpret_cast<const float2*>(a.center)[i], c2 = reinterpret_cast<const float2*>(a.center)[j];
    const float dx = __fsub_rn(c2.x, c1.x), dy = __fsub_rn(c2.y, c1.y);
    const double cd = sqrt((double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // math.sqrt of the fp32 sum
    // features in the model's order: center_distances_corrected, height_rel, iou, width_rel (matching_model_logreg.pckl)
    double d = __dmul_rn(cd, a.coef[0]);
    d = __dadd_rn(d, __dmul_rn((double)hrel, a.coef[1]));
    d = __dadd_rn(d, __dmul_rn((double)iou, a.coef[2]));
    d = __dadd_rn(d, __dmul_rn((double)wrel, a.coef[3]));
    d = __dadd_rn(d, a.intercept);
    double sc = __ddiv_rn(1.0, __dadd_rn(1.0, exp(-d)));                      // predict_proba[:, 1] = expit(decision)
    if (sc < a.clf_thr) return kReppInf;
    if (a.clf_mode == 0) sc = __dmul_rn(dot, sc);                             // 'dot'
    else if (a.clf_mode == 1) sc = __dmul_rn(__dmul_rn(si, sj), sc);          // 'max' (one-hot vectors: max = the score)
    else if (a.clf_mode == 2) sc = __dadd_rn(dot, sc);                        // 'dot_plus'
    return __dsub_rn(1.0, sc);                                                // 'raw' = 3: unchanged
}

__global__ void __launch_bounds__(kReppThreads) repp_link_kernel(const tscd_repp_link_args a) {
    __shared__ uint32_t used_r[kReppMaxDet / 32], used_c[kReppMaxDet / 32];
    __shared__ double red_d[kReppThreads / 32];
    __shared__ long long red_i[kReppThreads / 32];
    __shared__ int s_count, s_pairs;
    __shared__ long long s_best;
    const int f = blockIdx.x;                                   // links frame f -> f + 1
    const int o1 = a.frame_off[f], o2 = a.frame_off[f + 1], o3 = a.frame_off[f + 2];
    const int n1 = o2 - o1, n2 = o3 - o2;
    int32_t* pairs = a.pairs + (int64_t)f * a.max_det * 2;
    if (n1 <= 0 || n2 <= 0 || n1 > a.max_det || n2 > a.max_det || n1 > kReppMaxDet || n2 > kReppMaxDet) {
        if (threadIdx.x == 0) {
            a.pair_count[f] = 0;
            if (n1 > 0 && n2 > 0 && a.status) atomicMin(a.status, TSCD_ERR_CAPACITY);
        }
        return;
    }
    double* cd = a.ws_dist + (int64_t)f * a.ws_pitch;            // candidate distances
    int32_t* ci = a.ws_idx + (int64_t)f * a.ws_pitch;            // candidate row-major indices  i * n2 + j
    for (int w = threadIdx.x; w < kReppMaxDet / 32; w += blockDim.x) { used_r[w] = 0u; used_c[w] = 0u; }
    if (threadIdx.x == 0) { s_count = 0; s_pairs = 0; }
    __syncthreads();
    const int64_t total = (int64_t)n1 * n2;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
        const int i = (int)(e / n2), j = (int)(e - (int64_t)i * n2);
        const double d = repp_distance(a, o1 + i, o2 + j);
        if (d != kReppInf) {
            const int slot = atomicAdd(&s_count, 1);
            if (slot < a.ws_pitch) { cd[slot] = d; ci[slot] = (int)e; }
        }
    }
    __syncthreads();
    const int m = s_count;
    if (m > a.ws_pitch) {
        if (threadIdx.x == 0) { a.pair_count[f] = 0; if (a.status) atomicMin(a.status, TSCD_ERR_CAPACITY); }
        return;
    }
    // greedy: global minimum (ties: lowest row-major index) among candidates whose row and column are still free
    while (true) {
        double bd = kReppInf;
        long long bi = -1;
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const int e = ci[k];
            if (e < 0) continue;
            const int i = e / n2, j = e - i * n2;
            if (((used_r[i >> 5] >> (i & 31)) & 1u) || ((used_c[j >> 5] >> (j & 31)) & 1u)) { ci[k] = -1; continue; }
            const double d = cd[k];
            if (d < bd || (d == bd && (long long)e < bi) || bi < 0) { bd = d; bi = e; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { red_d[threadIdx.x >> 5] = bd; red_i[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kReppThreads / 32; ++w) {
                const double od = red_d[w];
                const long long oi = red_i[w];
                if (oi >= 0 && (bi < 0 || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
            }
            s_best = bi;
            if (bi >= 0) {
                const int i = (int)(bi / n2), j = (int)(bi - (long long)i * n2);
                used_r[i >> 5] |= 1u << (i & 31);
                used_c[j >> 5] |= 1u << (j & 31);
                pairs[2 * s_pairs] = i;
                pairs[2 * s_pairs + 1] = j;
                ++s_pairs;
            }
        }
        __syncthreads();
        if (s_best < 0) break;
    }
    if (threadIdx.x == 0) a.pair_count[f] = s_pairs;
}

}  // namespace tscd

extern "C" int tscd_repp_link(const tscd_repp_link_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->num_frames < 0 || a->max_det <= 0 || a->ws_pitch <= 0 || !a->frame_off || !a->bbox || !a->center || !a->score ||
        !a->cls || !a->pairs || !a->pair_count || !a->ws_dist || !a->ws_idx)
        return TSCD_ERR_INVALID_ARG;
    if (a->distance_func != 0 && a->distance_func != 1) return TSCD_ERR_INVALID_ARG;
    if (a->clf_mode < 0 || a->clf_mode > 3) return TSCD_ERR_INVALID_ARG;
    if (a->max_det > kReppMaxDet) return TSCD_ERR_CAPACITY;
    if (a->num_frames < 2) return TSCD_OK;
    repp_link_kernel<<<a->num_frames - 1, kReppThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}
