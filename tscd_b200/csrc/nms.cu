// K2: class-aware batched NMS, one CTA per frame (up to 4096 candidates; larger frames: csrc/nms_large.cu), bit-exact with
// torchvision's coordinate-trick path.
//
// Reference: torchvision.ops.batched_nms as called at yolox/models/tscd_head.py:1630 and
// yolox/models/post_process.py:58,73,510  (boxes.py `_batched_nms_coordinate_trick`):
//     offsets = class_id * (boxes.max() + 1);  keep = nms(boxes + offsets[:,None], scores, thr)
// nms = stable descending sort by score, greedy suppression where  inter/(Sa+Sb-inter) > thr.
// All arithmetic uses explicit round-to-nearest single-precision ops (no FMA contraction) so every IoU
// decision matches the CPU kernel; the comparison against `thr` is done in double like the CPU kernel.
//
// Algorithm (lazy greedy, no N^2 bitmask): candidates are sorted in shared memory (bitonic, 64-bit
// composite keys = score bits | inverted position, which reproduces the stable order).  They are then
// processed in chunks of 32: all threads test the chunk against the boxes kept so far (32 x kept IoUs) and build
// the 32x32 intra-chunk IoU bitmask, one warp resolves the chunk serially from the bitmask (ballot/shuffle).
// Later chunks are never touched before they are needed: work is O(kept * processed) instead of O(N^2) and stops
// as soon as max_keep boxes are kept (mode A only needs the first K=30 survivors of 750).
#include "common.cuh"
#include "nms.cuh"

namespace tscd {

constexpr int kNmsThreads = 256;

// -------------------------------------------------------------------------------------------------------------
// Per-class fast path, shared by both kernels.  After the coordinate-trick offset, boxes of different classes can
// only interact if the x-extents of their classes overlap (negative coordinates).  When no cross-class pair exceeds
// the IoU threshold -- the normal case -- greedy NMS decomposes EXACTLY into independent per-class problems: one warp per class walks
// the class's boxes in score order (lanes test the later members), all classes in parallel, and the keep list is
// the score-ordered compaction of the per-box flags.  Same offset boxes, same IoU predicate as the general paths.
// Returns false (nothing written) when the frame does not qualify: class ids outside [0, 256), a class with more
// than kClsMaxMembers boxes, or a cross-class pair above the IoU threshold; the caller then runs its general algorithm.
// -------------------------------------------------------------------------------------------------------------
constexpr int kClsMaxMembers = 128;

struct ClsScratch {          // carved from dynamic shared memory by the caller
    unsigned short* list;    // [n] sorted indices grouped by class (score order inside a class)
    unsigned char* flag;     // [n] 1 = kept
    unsigned char* scls;     // [n] class of sorted box r
    int* off;                // [257] class offsets into list
    float* lo;               // [256] min x1 of the class's offset boxes
    float* hi;               // [256] max x2
    int* misc;               // [4]
    unsigned char* ids;      // [256] non-empty class ids, ascending
};
__host__ __device__ inline size_t cls_scratch_bytes(int cap) { return (size_t)cap * 4 + 257 * 4 + 512 * 4 + 16 + 16 + 256; }
__device__ inline ClsScratch carve_cls_scratch(unsigned char* p, int cap) {
    ClsScratch c;
    c.off = reinterpret_cast<int*>(p); p += 257 * 4 + 12;
    c.lo = reinterpret_cast<float*>(p); p += 256 * 4;
    c.hi = reinterpret_cast<float*>(p); p += 256 * 4;
    c.misc = reinterpret_cast<int*>(p); p += 16;
    c.ids = p; p += 256;
    c.list = reinterpret_cast<unsigned short*>(p); p += (size_t)cap * 2;
    c.flag = p; p += cap;
    c.scls = p;
    return c;
}

// per-class greedy: lane j owns members j, j+32, ... (S slots, <= 32 * S members); the class leader walks the score order.
// S = 1 when every class holds at most 32 boxes (always the case for the stage's final NMS with <= 32 proposals per frame: a
// proposal contributes one row per class, post_process.py:36-46) -- a third of the IoU tests and no slot selection.
template <int S>
__device__ __forceinline__ void per_class_greedy(const tscd_nms_args& args, int ncl, const float4* sbox, const float* sarea, ClsScratch sc,
                                                 int warp, int lane, int nw) {
    const double thr = (double)args.iou_thresh;
    for (int ci = warp; ci < ncl; ci += nw) {
        const int c = sc.ids[ci];
        const int o0 = sc.off[c], cnt = sc.off[c + 1] - o0;
        float4 b[S];
        float ar[S];
        unsigned dead = 0;
#pragma unroll
        for (int q = 0; q < S; ++q) {
            const int k = q * 32 + lane;
            if (k < cnt) { const int r = sc.list[o0 + k]; b[q] = sbox[r]; ar[q] = sarea[r]; }
            else { b[q] = make_float4(0.f, 0.f, 0.f, 0.f); ar[q] = 0.f; dead |= 1u << q; }
        }
        for (int k = 0; k < cnt; ++k) {
            const int q0 = k >> 5, l0 = k & 31;
            // is member k still alive?  (owner lane l0, slot q0)
            const unsigned dk = __shfl_sync(0xffffffffu, dead, l0);
            if ((dk >> q0) & 1u) continue;
            float4 bk;
            float ak;
            {
                float4 src = b[0];
                float sa = ar[0];
#pragma unroll
                for (int q = 1; q < S; ++q)
                    if (q0 == q) { src = b[q]; sa = ar[q]; }
                bk.x = __shfl_sync(0xffffffffu, src.x, l0); bk.y = __shfl_sync(0xffffffffu, src.y, l0);
                bk.z = __shfl_sync(0xffffffffu, src.z, l0); bk.w = __shfl_sync(0xffffffffu, src.w, l0);
                ak = __shfl_sync(0xffffffffu, sa, l0);
            }
#pragma unroll
            for (int q = 0; q < S; ++q) {
                const int m = q * 32 + lane;
                if (m > k && m < cnt && !((dead >> q) & 1u) && iou_gt(bk, ak, b[q], ar[q], thr)) dead |= 1u << q;
            }
        }
#pragma unroll
        for (int q = 0; q < S; ++q) {
            const int k = q * 32 + lane;
            if (k < cnt) sc.flag[sc.list[o0 + k]] = ((dead >> q) & 1u) ? 0 : 1;
        }
    }
}

// skey: sorted keys (low 32 bits = inverted original position); sbox/sarea: sorted offset boxes.
__device__ bool nms_per_class(const tscd_nms_args& args, int frame, int n, const unsigned long long* skey, const float4* sbox,
                              const float* sarea, const int32_t* gcls, ClsScratch sc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) { sc.misc[0] = 0; sc.misc[1] = 0; }
    for (int c = tid; c < 257; c += blockDim.x) sc.off[c] = 0;
    __syncthreads();
    // class of every sorted box, class histogram
    int bad = 0;
    for (int r = tid; r < n; r += blockDim.x) {
        const int pos = (int)(0xffffffffu - (uint32_t)(skey[r] & 0xffffffffull));
        const int c = gcls[pos];
        if (c < 0 || c > 255) { bad = 1; continue; }
        sc.scls[r] = (unsigned char)c;
        atomicAdd(&sc.off[c + 1], 1);
    }
    if (bad) sc.misc[0] = 1;
    __syncthreads();
    if (sc.misc[0]) return false;
    if (warp == 0) {                       // exclusive prefix over the 256 class counts (8 per lane), size check
        int loc[8], tot = 0, big = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { loc[j] = sc.off[1 + lane * 8 + j]; tot += loc[j]; big |= loc[j] > kClsMaxMembers; }
        int inc = warp_incl_scan(tot, lane);
        int run = inc - tot;
#pragma unroll
        for (int j = 0; j < 8; ++j) { run += loc[j]; sc.off[1 + lane * 8 + j] = run; }
        __syncwarp();
        if (__any_sync(0xffffffffu, big) && lane == 0) sc.misc[0] = 1;
        {
            int mxc = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) mxc = max(mxc, loc[j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mxc = max(mxc, __shfl_xor_sync(0xffffffffu, mxc, o));
            if (lane == 0) sc.misc[3] = mxc;
        }
        // compact list of the non-empty classes (ascending ids), kept in sc.ids
        int ncl = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j * 32 + lane;
            const bool ne = sc.off[c + 1] - sc.off[c] > 0;      // prefix values of the lower classes are final: same warp, in order
            __syncwarp();
            const unsigned bal = __ballot_sync(0xffffffffu, ne);
            if (ne) sc.ids[ncl + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)c;
            ncl += __popc(bal);
        }
        if (lane == 0) sc.misc[2] = ncl;
    }
    __syncthreads();
    if (sc.misc[0]) return false;
    const int ncl = sc.misc[2];
    // class lists in score order + class x-bands: warp per class, ballot compaction over the sorted boxes
    for (int ci = warp; ci < ncl; ci += nw) {
        const int c = sc.ids[ci];
        const int o0 = sc.off[c], cnt = sc.off[c + 1] - o0;
        int filled = 0;
        float lo = INFINITY, hi = -INFINITY;
        for (int r0 = 0; r0 < n && filled < cnt; r0 += 32) {
            const int r = r0 + lane;
            const bool m = r < n && sc.scls[r] == c;
            const unsigned bal = __ballot_sync(0xffffffffu, m);
            if (m) {
                sc.list[o0 + filled + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)r;
                lo = fminf(lo, sbox[r].x); hi = fmaxf(hi, sbox[r].z);
            }
            filled += __popc(bal);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if (lane == 0) { sc.lo[c] = lo; sc.hi[c] = hi; }
    }
    __syncthreads();
    // Classes whose x-bands overlap (negative coordinates near the image border) may interact.  The decomposition stays
    // exact as long as no CROSS-class pair actually exceeds the IoU threshold: test exactly those pairs (warp per class,
    // lanes over the member pairs of every overlapping later class); a single hit sends the frame to the general path.
    {
        const double thr_x = (double)args.iou_thresh;
        for (int ci = warp; ci < ncl; ci += nw) {
            const int c = sc.ids[ci];
            const int oc = sc.off[c], nc = sc.off[c + 1] - oc;
            const float l = sc.lo[c], h = sc.hi[c];
            for (int di = ci + 1; di < ncl; ++di) {
                const int d = sc.ids[di];
                if (!(fminf(h, sc.hi[d]) > fmaxf(l, sc.lo[d]))) continue;
                const int od = sc.off[d], nd = sc.off[d + 1] - od;
                bool hit = false;
                for (int ii = 0; ii < nc; ++ii) {              // lanes over the members of d, one member of c at a time
                    const int i = sc.list[oc + ii];
                    const float4 bi = sbox[i];
                    if (!(bi.z > sc.lo[d])) continue;           // box i does not reach into the band of class d (d > c: further right)
                    const float si = sarea[i];
                    for (int jj = lane; jj < nd; jj += 32) {
                        const int j = sc.list[od + jj];
                        hit = hit || iou_gt(bi, si, sbox[j], sarea[j], thr_x);
                    }
                }
                if (__any_sync(0xffffffffu, hit)) { if (lane == 0) sc.misc[0] = 1; }
            }
        }
    }
    __syncthreads();
    if (sc.misc[0]) return false;

    if (sc.misc[3] <= 32) per_class_greedy<1>(args, ncl, sbox, sarea, sc, warp, lane, nw);
    else if (sc.misc[3] <= 64) per_class_greedy<2>(args, ncl, sbox, sarea, sc, warp, lane, nw);
    else per_class_greedy<kClsMaxMembers / 32>(args, ncl, sbox, sarea, sc, warp, lane, nw);
    __syncthreads();
    // score-ordered compaction of the kept boxes, truncated to max_keep
    {
        const int max_keep = args.max_keep;
        const int lim = max_keep + (args.strict_keep ? 1 : 0);     // strict: look one survivor further to detect the overflow
        int32_t* keep = args.keep + (int64_t)frame * max_keep;
        int* carry = &sc.misc[1];
        int* scan = sc.off;              // class offsets are no longer needed: reuse as scan scratch (>= 33 ints)
        __syncthreads();
        for (int r0 = 0; r0 < n; r0 += blockDim.x) {
            const int r = r0 + tid;
            const int f = (r < n && sc.flag[r]) ? 1 : 0;
            int tot;
            const int ex = block_excl_scan(f, scan, &tot);
            const int base = *carry;
            if (f && base + ex < max_keep) keep[base + ex] = (int)(0xffffffffu - (uint32_t)(skey[r] & 0xffffffffull));
            __syncthreads();
            if (tid == 0) *carry = base + tot;
            __syncthreads();
            if (*carry >= lim) break;
        }
        if (tid == 0) {
            args.keep_count[frame] = min(*carry, max_keep);
            if (args.strict_keep && *carry > max_keep) atomicMin(args.status, TSCD_ERR_CAPACITY);
        }
    }
    return true;
}

// only_redo: second launch after nms_matrix_kernel -- only the frames it marked with keep_count == -1 (per-class decomposition
// not applicable, too large for the suppression matrix) are processed
__global__ void __launch_bounds__(kNmsThreads, 6) nms_kernel(const tscd_nms_args args, int smem_cap, int only_redo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int frame = blockIdx.x;
    if (only_redo && args.keep_count[frame] != -1) return;
    int n = args.count[frame];
    if (n > smem_cap) {
        if (threadIdx.x == 0) { atomicMin(args.status, TSCD_ERR_CAPACITY); args.keep_count[frame] = 0; }
        return;
    }
    if (n <= 0) {
        if (threadIdx.x == 0) args.keep_count[frame] = 0;
        return;
    }
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem_raw);      // [smem_cap]
    float4* sbox = reinterpret_cast<float4*>(skey + smem_cap);                        // [n] sorted, offset boxes
    float* sarea = reinterpret_cast<float*>(sbox + smem_cap);                         // [n]
    unsigned char* dead = reinterpret_cast<unsigned char*>(sarea + smem_cap);         // [n] ints: positions of the kept boxes
    __shared__ unsigned int cmask[32];
    __shared__ float red[kNmsThreads / 32];
    __shared__ int s_nkept;
    __shared__ unsigned int s_deadbits;
    __shared__ SelSmem rs;

    const int64_t base = (int64_t)frame * args.cand_cap;
    const float* gscore = args.score + base;
    const float4* gbox = reinterpret_cast<const float4*>(args.box) + base;
    const int32_t* gcls = args.cls + base;
    // optional tie-break keys (tscd_select's cand_rank: candidates in anchor order, rank = objectness key << 16 | 0xffff - position):
    // equal scores are ordered by descending rank, the position is the rank's low half
    const uint32_t* grank = args.rank ? args.rank + base : nullptr;
    auto low_of = [&](int i) -> uint32_t { return grank ? grank[i] : 0xffffffffu - (uint32_t)i; };
    auto pos_of = [&](unsigned long long key) -> int {
        const uint32_t lo32 = (uint32_t)(key & 0xffffffffull);
        return grank ? (int)(0xffffu - (lo32 & 0xffffu)) : (int)(0xffffffffu - lo32);
    };

    // ---- max coordinate (over ALL boxes: the coordinate-trick offset unit) --------------------------------
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float4 b = gbox[i];
        mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    }
    mx = warp_maxf(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < kNmsThreads / 32; ++w) mx = fmaxf(mx, red[w]);
    const float off_unit = __fadd_rn(mx, 1.f);  // boxes.max() + 1
    const double thr = (double)args.iou_thresh;
    const int max_keep = args.max_keep;
    const int lim = max_keep + (args.strict_keep ? 1 : 0);         // strict: look one survivor further to detect the overflow
    int32_t* keep = args.keep + (int64_t)frame * max_keep;
    int* s_keptidx = reinterpret_cast<int*>(dead);     // sorted positions of the boxes kept so far (<= min(n, max_keep))

    // Top-K use (mode A keeps the first 30 survivors of 750): sorting all candidates is three quarters of this kernel.
    // Attempt 0 takes only the candidates whose score reaches the kPartial-th largest one (block radix select; every
    // tie at the threshold is included, so the subset is exactly a PREFIX of the fully sorted order), sorts those 256 at
    // most and runs the same greedy loop; if that prefix is exhausted before max_keep boxes survive, attempt 1 redoes
    // the frame with the full sort.
    constexpr int kPartial = 128, kPartialCap = 256;
    const bool try_partial = max_keep * 4 <= kPartial && n > kPartialCap;
    for (int attempt = try_partial ? 0 : 1; attempt < 2; ++attempt) {
        int n_work = n;                          // number of sorted candidates this attempt can look at
        if (attempt == 0) {
            uint32_t* k32 = reinterpret_cast<uint32_t*>(sarea);      // scratch: sarea is written later
            for (int i = threadIdx.x; i < n; i += blockDim.x) k32[i] = f2ord(gscore[i]);
            __syncthreads();
            uint32_t T;
            int req;
            radix_select_kth<uint32_t>(k32, n, kPartial, &rs, &T, &req);
            int cnt = 0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) cnt += k32[i] >= T ? 1 : 0;
            int tot;
            block_excl_scan(cnt, rs.scan, &tot);
            if (tot > kPartialCap) continue;                              // too many ties at the threshold: full path
            // positions in ascending order (any order works: the sort key carries the position)
            if (threadIdx.x == 0) rs.misc[3] = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                if (k32[i] >= T) {
                    const int slot = atomicAdd(&rs.misc[3], 1);
                    skey[slot] = ((unsigned long long)k32[i] << 32) | (unsigned long long)low_of(i);
                }
            }
            __syncthreads();
            n_work = tot;
            block_sort_desc64_dyn<unsigned long long>(skey, n_work, kPartialCap);
        } else {
            if (n <= 32) {
                // tiny frames (the unrefined "ori" rows: <= 30 candidates): one warp sorts in registers
                if (threadIdx.x < 32) {
                    const int lane = threadIdx.x;
                    unsigned long long v = lane < n ? (((unsigned long long)f2ord(gscore[lane]) << 32) | (unsigned long long)low_of(lane)) : 0ull;
                    for (int k = 2; k <= 32; k <<= 1)
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
                            const bool keep_max = (((lane & j) == 0) == ((lane & k) == 0));
                            v = keep_max ? (v > o ? v : o) : (v < o ? v : o);
                        }
                    skey[lane] = v;
                }
                __syncthreads();
            } else {
                for (int i = threadIdx.x; i < n; i += blockDim.x)
                    skey[i] = ((unsigned long long)f2ord(gscore[i]) << 32) | (unsigned long long)low_of(i);
                __syncthreads();
                block_sort_desc64_dyn<unsigned long long>(skey, n, smem_cap);
            }
        }

        for (int r = threadIdx.x; r < n_work; r += blockDim.x) {
            int pos = pos_of(skey[r]);
            float4 b = gbox[pos];
            float off = __fmul_rn((float)gcls[pos], off_unit);
            b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off);
            b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
            sbox[r] = b;
            sarea[r] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
        }
        if (threadIdx.x == 0) s_nkept = 0;
        __syncthreads();

        for (int c0 = 0; c0 < n_work; c0 += 32) {
            const int cn = min(32, n_work - c0);
            // (0) fully lazy: only THIS chunk is tested against the boxes kept so far (32 x kept IoUs); later chunks are
            //     never touched if max_keep is reached first (mode A wants the first 30 survivors of 750)
            if (threadIdx.x < 32) cmask[threadIdx.x] = 0u;
            if (threadIdx.x == 0) s_deadbits = 0u;
            __syncthreads();
            const int nk0 = s_nkept;
            {
                unsigned mydead = 0u;
                const int l = threadIdx.x & 31;
                if (l < cn) {
                    const float4 bl = sbox[c0 + l];
                    const float sl = sarea[c0 + l];
                    for (int k = threadIdx.x >> 5; k < nk0 && !mydead; k += kNmsThreads / 32) {
                        const int i = s_keptidx[k];
                        if (iou_gt(sbox[i], sarea[i], bl, sl, thr)) mydead = 1u << l;
                    }
                }
                if (mydead) atomicOr(&s_deadbits, mydead);
            }
            // (1) intra-chunk bitmask: bit j of cmask[l] set if earlier lane j suppresses lane l
            for (int pr = threadIdx.x; pr < 32 * 32; pr += blockDim.x) {
                int l = pr >> 5, j = pr & 31;
                if (j < l && l < cn) {
                    if (iou_gt(sbox[c0 + j], sarea[c0 + j], sbox[c0 + l], sarea[c0 + l], thr)) atomicOr(&cmask[l], 1u << j);
                }
            }
            __syncthreads();
            // (2) serial resolve by warp 0
            if (threadIdx.x < 32) {
                const int lane = threadIdx.x;
                bool alive = (lane < cn) && !((s_deadbits >> lane) & 1u);
                unsigned alive_bits = __ballot_sync(0xffffffffu, alive);
                unsigned my = cmask[lane];
                unsigned kept = 0u;
    #pragma unroll
                for (int l = 0; l < 32; ++l) {
                    unsigned m = __shfl_sync(0xffffffffu, my, l);
                    if (((alive_bits >> l) & 1u) && !(m & kept)) kept |= 1u << l;
                }
                int nk = nk0;
                // truncate to max_keep
                int rank = __popc(kept & ((1u << lane) - 1u));
                bool mine = (kept >> lane) & 1u;
                if (mine && nk + rank < lim) {
                    if (nk + rank < max_keep) keep[nk + rank] = pos_of(skey[c0 + lane]);
                    s_keptidx[nk + rank] = c0 + lane;
                }
                if (lane == 0) s_nkept = nk + min(__popc(kept), lim - nk);
            }
            __syncthreads();
            if (s_nkept >= lim) break;
        }
        __syncthreads();
        if (s_nkept >= lim || n_work == n) break;     // done; otherwise the prefix was too short: full sort
    }
    if (threadIdx.x == 0) {
        args.keep_count[frame] = min(s_nkept, max_keep);
        if (args.strict_keep && s_nkept > max_keep) atomicMin(args.status, TSCD_ERR_CAPACITY);
    }
}

// -------------------------------------------------------------------------------------------------------------
// Suppression-matrix variant for frames that keep most of their candidates (the final per-class NMS keeps up to all
// n <= 768 boxes, so the lazy kernel's O(kept * N) apply phase degenerates to N^2 with a barrier-heavy chunk loop).
// Same sort, same offset boxes, same IoU predicate; then
//   (1) all threads build the upper-triangular bit matrix  M[i][w] bit b = box i suppresses box 32w+b (> i),
//   (2) one warp walks the sorted boxes 32 at a time: the diagonal block resolves the word with shuffles, the rows of
//       the newly kept boxes are OR-ed into the per-lane `removed` words.
// -------------------------------------------------------------------------------------------------------------
constexpr int kNmsMatThreads = 512;
constexpr int kNmsMatCap = 768;      // 768 x 24 words = 72 KB
constexpr int kNmsMatInline = 256;   // the matrix shares the per-class kernel only up to 256 rows (8 KB)

__global__ void __launch_bounds__(kNmsMatThreads, 3) nms_matrix_kernel(const tscd_nms_args args, int smem_cap, int has_matrix) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int frame = blockIdx.x;
    int n = args.count[frame];
    if (n > smem_cap) {
        if (threadIdx.x == 0) { atomicMin(args.status, TSCD_ERR_CAPACITY); args.keep_count[frame] = 0; }
        return;
    }
    if (n <= 0) {
        if (threadIdx.x == 0) args.keep_count[frame] = 0;
        return;
    }
    int n64 = 1;
    while (n64 < n) n64 <<= 1;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem_raw);      // [n64]
    float4* sbox = reinterpret_cast<float4*>(skey + smem_cap);                        // [n] sorted, offset boxes
    float* sarea = reinterpret_cast<float*>(sbox + smem_cap);                         // [n]
    uint32_t* smask = reinterpret_cast<uint32_t*>(sarea + smem_cap);                  // [n][W]
    __shared__ float red[kNmsMatThreads / 32];

    const int64_t base = (int64_t)frame * args.cand_cap;
    const float* gscore = args.score + base;
    const float4* gbox = reinterpret_cast<const float4*>(args.box) + base;
    const int32_t* gcls = args.cls + base;

    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n64; i += blockDim.x) {
        unsigned long long v = 0ull;
        if (i < n) {
            v = ((unsigned long long)f2ord(gscore[i]) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
            float4 b = gbox[i];
            mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
        }
        skey[i] = v;
    }
    mx = warp_maxf(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < kNmsMatThreads / 32; ++w) mx = fmaxf(mx, red[w]);
    const float off_unit = __fadd_rn(mx, 1.f);  // boxes.max() + 1

    block_sort_desc64_dyn<unsigned long long>(skey, n, smem_cap);

    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        int pos = (int)(0xffffffffu - (uint32_t)(skey[r] & 0xffffffffull));
        float4 b = gbox[pos];
        float off = __fmul_rn((float)gcls[pos], off_unit);
        b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off);
        b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
        sbox[r] = b;
        sarea[r] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    }
    __syncthreads();
    // the suppression-matrix area doubles as scratch of the per-class fast path
    if (nms_per_class(args, frame, n, skey, sbox, sarea, gcls, carve_cls_scratch(reinterpret_cast<unsigned char*>(smask), smem_cap)))
        return;
    if (!has_matrix) {                  // no room for the suppression matrix: hand the frame to the lazy kernel (second launch)
        if (threadIdx.x == 0) args.keep_count[frame] = -1;
        return;
    }

    const double thr = (double)args.iou_thresh;
    const int W = (n + 31) >> 5;
    // (1) suppression matrix; rows are dealt in a snake order so that long and short rows alternate per thread
    const int T = blockDim.x;
    for (int k = 0;; ++k) {
        const int i = (k & 1) ? (T * (k + 1) - 1 - (int)threadIdx.x) : (T * k + (int)threadIdx.x);
        if (T * k >= n) break;
        if (i >= n) continue;
        const float4 bi = sbox[i];
        const float si = sarea[i];
        for (int w = i >> 5; w < W; ++w) {
            uint32_t bits = 0u;
            const int j0 = w << 5;
            const int jlo = max(i + 1, j0), jhi = min(n, j0 + 32);
            for (int j = jlo; j < jhi; ++j) {
                const float4 bj = sbox[j];
                // cheap x-extent rejection first: boxes of different classes sit (max+1) apart after the offset
                if (fminf(bi.z, bj.z) > fmaxf(bi.x, bj.x)) {
                    if (iou_gt(bi, si, bj, sarea[j], thr)) bits |= 1u << (j - j0);
                }
            }
            smask[i * W + w] = bits;
        }
    }
    __syncthreads();

    // (2) resolve, 32 sorted boxes per step
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int max_keep = args.max_keep;
        const int lim = max_keep + (args.strict_keep ? 1 : 0);
        int32_t* keep = args.keep + (int64_t)frame * max_keep;
        uint32_t removed = 0u;           // lane w < W: suppressed flags of boxes [32w, 32w+32)
        int nk = 0;
        for (int w = 0; w < W && nk < lim; ++w) {
            const int i = (w << 5) + lane;
            const uint32_t rem_w = __shfl_sync(0xffffffffu, removed, w);
            const uint32_t diag = i < n ? smask[i * W + w] : 0u;          // boxes of this word that box i suppresses
            uint32_t alive = ~rem_w;
            if (n - (w << 5) < 32) alive &= (1u << (n - (w << 5))) - 1u;
            uint32_t kept = 0u;
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const uint32_t dl = __shfl_sync(0xffffffffu, diag, l);
                if ((alive >> l) & 1u) { kept |= 1u << l; alive &= ~dl; }
            }
            // truncate to max_keep, emit, and apply the kept rows to the later words
            const int room = lim - nk;
            const int rank = __popc(kept & ((1u << lane) - 1u));
            const bool mine = ((kept >> lane) & 1u) && nk + rank < max_keep;
            if (mine) keep[nk + rank] = (int)(0xffffffffu - (uint32_t)(skey[i] & 0xffffffffull));
            uint32_t kk = kept;
            while (kk) {
                const int l = __ffs(kk) - 1;
                kk &= kk - 1;
                if (lane > w && lane < W) removed |= smask[((w << 5) + l) * W + lane];
            }
            nk += min(__popc(kept), room);
        }
        if (lane == 0) {
            args.keep_count[frame] = min(nk, max_keep);
            if (args.strict_keep && nk > max_keep) atomicMin(args.status, TSCD_ERR_CAPACITY);
        }
    }
}

}  // namespace tscd

extern "C" int tscd_nms(const tscd_nms_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->num_frames < 0 || a->cand_cap <= 0 || a->max_keep <= 0 || !(a->iou_thresh >= 0.f)) return TSCD_ERR_INVALID_ARG;
    if (a->num_frames == 0) return TSCD_OK;
    if (a->rank && (a->cand_cap > 65535 || a->cand_cap <= 64 || (int64_t)a->max_keep * 4 >= a->cand_cap)) return TSCD_ERR_UNSUPPORTED;  // top-K kernel only
    if (a->cand_cap > kNmsCap) return nms_large_launch(*a, reinterpret_cast<cudaStream_t>(stream));
    int cap = a->cand_cap;
    int cap64 = 1;
    while (cap64 < cap) cap64 <<= 1;   // sort buffer must hold the padded power of two
    if (cap64 < kNmsMatThreads) cap64 = kNmsMatThreads;   // ... and E * blockDim.x keys of the block sort (either kernel)
    size_t smem = (size_t)cap64 * (8 + 16 + 4 + 4) + 16;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int only_redo = 0;
    if (cap > 64 && (int64_t)a->max_keep * 4 >= cap && !a->rank) {
        // most candidates survive: per-class decomposition, then (cap <= 768) the suppression-matrix algorithm in the same
        // kernel or (larger caps) the lazy kernel for the frames the decomposition does not cover.  (The lazy kernel stops
        // early and wins when only the first few survivors are wanted, and for tiny frames.)
        // The suppression matrix (frames the per-class decomposition does not cover: cross-class overlaps, which the coordinate
        // trick's class offsets make rare) stays in this kernel only while it is small: 72 KB of it for 750 rows would cap the
        // kernel at two CTAs per SM for every frame -- larger frames mark themselves and take the second launch instead.
        const bool inline_matrix = cap <= kNmsMatInline;
        size_t smem_m = inline_matrix ? (size_t)cap * ((cap + 31) / 32) * 4 : 0;
        if (smem_m < cls_scratch_bytes(cap64)) smem_m = cls_scratch_bytes(cap64);
        smem_m += (size_t)cap64 * (8 + 16 + 4) + 16;
        if (cudaFuncSetAttribute(nms_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m) != cudaSuccess)
            return TSCD_ERR_CUDA;
        nms_matrix_kernel<<<a->num_frames, kNmsMatThreads, smem_m, st>>>(*a, cap64, inline_matrix ? 1 : 0);
        TSCD_CUDA_CHECK_LAUNCH();
        if (inline_matrix) return TSCD_OK;
        only_redo = 1;
    }
    if (cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return TSCD_ERR_CUDA;
    nms_kernel<<<a->num_frames, kNmsThreads, smem, st>>>(*a, cap64, only_redo);
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}
