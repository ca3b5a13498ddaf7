// K1 over the FUSED head layout: one 2^k-byte row per anchor, [reg 4 | obj 1 | cls C | pad] in fp16 (64 B for C <= 27,
// 128 B for C <= 59), plus the objectness logits once more as a dense [F, A] plane.  This is the layout the drop-in head
// emits (tscd_pack_head / tscd_b200.head) and the one forward_host reads in place out of pinned host memory: an anchor's
// whole head output is ONE aligned 64-byte row, so a survivor costs one full-sector read instead of a 50-byte class row and
// an 8-byte regression row fetched as separate bursts.
//
// Reference semantics: postpro_woclass (yolox/models/post_process.py:464-521): topk(obj, P) -> rows
// [box4, obj, class_conf, class_pred] in descending objectness order (ties: lower anchor id first, the documented rule of
// include/tscd_b200.h), score = obj * class_conf.
//
// Mode A, one CTA per frame:
//   1. objectness plane -> one 16-bit order-preserving key per anchor in shared memory.  sigmoid is monotone, so the order
//      of the fp32 scores is the order of the fp16 logits EXCEPT where several logits share one fp32 score (saturation:
//      x > 9 or x < -80, and |x| < 2^-9 where fp16 steps are finer than the fp32 score's); those keys are replaced by the lowest key of their plateau (binary search on the
//      same sigmoid the scores use), which makes "order by key" identical to "order by score" -- no fp32 score of the 6804
//      anchors is ever computed.  tests/test_gpu_selection.py checks the equivalence exhaustively over all 65536 fp16 values.
//   2. two-digit radix select of the P-th largest key: the high-byte histogram is built WHILE the keys are produced, into
//      per-warp private 16-bit histograms (sigmoid logits share a few exponent bytes: a shared histogram serialises the CTA on a
//      handful of bins); the low-byte pass only touches the few hundred keys of the threshold's high byte.  Stable compaction:
//      every thread owns a contiguous run of keys, counts (> T, == T) locally, ONE block scan of the packed counts places them.
//   3. every survivor's row is requested with cp.async (16-byte global -> shared copies, 4 or 8 adjacent lanes per row = one
//      coalesced 64 / 128-byte request) as soon as its position is known.
//   4. (only without cand_rank) block bitonic sort on 32-bit keys; then one thread per survivor builds the record from the
//      staged row: class max / arg-max, the two sigmoids, box decode.
#include "common.cuh"

namespace tscd {

constexpr int kRowsThreads = 512;

__device__ __forceinline__ uint32_t ord16(uint32_t hb) { return (hb & 0x8000u) ? (~hb & 0xffffu) : (hb | 0x8000u); }
__device__ __forceinline__ uint32_t unord16(uint32_t k) { return (k & 0x8000u) ? (k & 0x7fffu) : (~k & 0xffffu); }
__device__ __forceinline__ float hbits2f(uint32_t hb) { return __half2float(__ushort_as_half((unsigned short)hb)); }

// Order-preserving 16-bit key of an fp16 value such that  key(x) < key(y)  <=>  score(x) < score(y)  where score is the fp32
// value the reference compares: sigmoidf_ref(x) for logits (sig), x itself otherwise.
__device__ __forceinline__ uint32_t canon_key16(uint32_t hb, bool sig) {
    if (hb == 0x8000u) hb = 0u;                         // -0 and +0 are the same score
    uint32_t k = ord16(hb);
    if (!sig) return k;
    const float x = hbits2f(hb);
    // plateaus of the fp32 sigmoid over fp16 inputs: saturation (x > 9: 1 - s below the fp32 spacing; x < -80: denormal /
    // zero scores) and the neighbourhood of 0 (fp16 steps of 2^-24 .. 2^-20 move s = 0.5 + x/4 by less than one fp32 ulp).
    // There the key becomes the LOWEST key with the same score (binary search on the sigmoid the scores use; monotone).
    const bool hi_sat = x > 9.f, lo_sat = x < -80.f, near0 = fabsf(x) < 0.001953125f;
    if (hi_sat || lo_sat || near0) {
        const float s = sigmoidf_ref(x);
        uint32_t lo = hi_sat ? ord16(0x4880u) /* 9.0 */ : (lo_sat ? ord16(0xfc00u) /* -inf */ : ord16(0x9800u) /* -2^-9 */), hi = k;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sigmoidf_ref(hbits2f(unord16(mid))) == s) hi = mid; else lo = mid + 1;
        }
        k = lo;
    }
    return k;
}

__global__ void canon_key_debug_kernel(const unsigned short* hb, int n, int sig, unsigned short* key, float* score) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = (unsigned short)canon_key16(hb[i], sig != 0);
    const float x = hbits2f(hb[i]);
    score[i] = sig ? sigmoidf_ref(x) : x;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// RPV = 16-byte vectors per fused row (4: 64-byte rows, 8: 128-byte rows)
template <int RPV>
__global__ void __launch_bounds__(kRowsThreads, RPV == 4 ? 3 : 2)
select_rows_kernel(const tscd_select_args args, int sort_cap, int take_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = kRowsThreads, NW = NT / 32;
    const int frame = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const tscd_anchors& an = args.anchors;
    const int A = an.level_start[an.num_levels];
    const int C = args.num_classes;
    const bool sig = args.apply_sigmoid != 0;
    const int A8 = (A + 15) & ~7;                                          // room for the alignment shift of the keys

    // the row staging area doubles as the per-warp digit histograms of the first radix pass (never less than NW * 512 bytes)
    const size_t rows_bytes = max((size_t)take_cap * RPV * 16, (size_t)NW * 512);
    uint4* rows = reinterpret_cast<uint4*>(smem_raw);                      // [take_cap][RPV] staged survivor rows
    uint32_t* sort32 = reinterpret_cast<uint32_t*>(smem_raw + rows_bytes); // [sort_cap] (key << 16 | 0xffff - position)
    unsigned short* k16_base = reinterpret_cast<unsigned short*>(sort32 + sort_cap); // [A8] keys
    unsigned short* sel16 = k16_base + A8;                                 // [take_cap] anchor id of compaction position i
    __shared__ SelSmem s;

    // ---- 1. objectness plane -> keys, and the histogram of their high bytes --------------------------------------
    // Sigmoid logits share a few exponent bytes, so one shared histogram would serialise the whole CTA on a handful of bins:
    // every warp counts into its OWN 256 x 16-bit histogram (two bins per word, plain shared atomics, no cross-warp
    // contention; a warp sees < 65536 keys), merged once afterwards.
    uint32_t* whist = reinterpret_cast<uint32_t*>(smem_raw) + wid * 128;
    for (int i = tid; i < NW * 128; i += NT) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
    __syncthreads();
    const __half* op = reinterpret_cast<const __half*>(args.obj.ptr[0]) + (int64_t)frame * args.obj.frame_stride[0];
    const int head = min(A, (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(op) & 15u)) & 15u) >> 1));   // elements before 16-byte alignment
    const int sh = (8 - head) & 7;
    unsigned short* k16 = k16_base + sh;                                   // k16[head + 8 g] is 16-byte aligned in shared memory
    const int nvec = (A - head) >> 3;
    auto count_hi = [&](uint32_t key) { atomicAdd(&whist[key >> 9], 1u << ((key >> 4) & 16u)); };   // bin = key >> 8: word bin / 2, half bin & 1
    for (int g = tid; g < nvec; g += NT) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(op + head) + g);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t k0 = canon_key16(w[q] & 0xffffu, sig), k1 = canon_key16(w[q] >> 16, sig);
            count_hi(k0);
            count_hi(k1);
            o[q] = k0 | (k1 << 16);
        }
        *reinterpret_cast<uint4*>(k16 + head + 8 * g) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    for (int i = tid; i < head + (A - head - 8 * nvec); i += NT) {         // unaligned head + tail
        const int a = i < head ? i : 8 * nvec + i;
        const uint32_t k = canon_key16(__half_as_ushort(__ldg(op + a)), sig);
        count_hi(k);
        k16[a] = (unsigned short)k;
    }
    __syncthreads();

    // ---- 2. P-th largest key (two 8-bit digits), stable compaction in ascending anchor order --------------------------
    const int take_k = min(min(args.pre_k, A), take_cap);
    // suffix sums over the 256 bins from the top digit down (warp 0: lane l owns digits [8l, 8l+8)): the digit holding the
    // `remaining`-th largest key -> s.misc[0], its rank inside that digit -> s.misc[1]
    auto find_digit = [&](int remaining) {
        if (tid < 32) {
            int loc[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = s.hist[255 - (lane * 8 + j)]; tot += loc[j]; }
            const int inc = warp_incl_scan(tot, lane);
            int cum = inc - tot;
            if (cum < remaining && remaining <= inc) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (cum < remaining && remaining <= cum + loc[j]) { s.misc[0] = 255 - (lane * 8 + j); s.misc[1] = remaining - cum; }
                    cum += loc[j];
                }
            }
        }
    };
    if (tid < 256) {                                                       // merge the per-warp histograms
        const uint32_t* h = reinterpret_cast<const uint32_t*>(smem_raw) + (tid >> 1);
        const int hs = 16 * (tid & 1);
        int tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += (int)((h[w * 128] >> hs) & 0xffffu);
        s.hist[tid] = tot;
    }
    __syncthreads();
    find_digit(take_k);
    __syncthreads();
    const uint32_t d_hi = (uint32_t)s.misc[0];
    const int rem_hi = s.misc[1];
    __syncthreads();
    if (tid < 256) s.hist[tid] = 0;
    __syncthreads();
    // low byte among the keys of the threshold's high byte (a few hundred keys spread over 256 bins)
    for (int g = tid; g < nvec; g += NT) {
        const uint4 kv = *reinterpret_cast<const uint4*>(k16 + head + 8 * g);
        const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (((w[q] >> 8) & 255u) == d_hi) atomicAdd(&s.hist[w[q] & 255u], 1);
            if ((w[q] >> 24) == d_hi) atomicAdd(&s.hist[(w[q] >> 16) & 255u], 1);
        }
    }
    for (int i = tid; i < head + (A - head - 8 * nvec); i += NT) {
        const uint32_t k = k16[i < head ? i : 8 * nvec + i];
        if ((k >> 8) == d_hi) atomicAdd(&s.hist[k & 255u], 1);
    }
    __syncthreads();
    find_digit(rem_hi);
    __syncthreads();
    const uint32_t Tk = (d_hi << 8) | (uint32_t)s.misc[0];
    const int r_eq = s.misc[1];                                            // keys equal to Tk that belong to the top take_k
    // compaction: a thread owns a CONTIGUOUS run of keys (whole aligned 16-byte vectors of the shifted key array), counts its
    // keys above / equal to the threshold, one block scan of the packed counts gives its first output position
    const int V = (A + sh + 7) >> 3, VPT = (V + NT - 1) / NT;
    const int v_lo = min(V, tid * VPT), v_hi = min(V, v_lo + VPT);
    uint32_t cnt = 0;                                                      // keys > Tk | keys == Tk << 16
    for (int v = v_lo; v < v_hi; ++v) {
        const uint4 kv = *reinterpret_cast<const uint4*>(k16_base + 8 * v);
        const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t u = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
            if ((unsigned)(8 * v + e - sh) < (unsigned)A) cnt += (u > Tk ? 1u : 0u) + (u == Tk ? 0x10000u : 0u);
        }
    }
    int tot_unused;
    const uint32_t before = (uint32_t)block_excl_scan((int)cnt, s.scan, &tot_unused);
    int eq_idx = (int)(before >> 16);
    int out = (int)(before & 0xffffu) + min(eq_idx, r_eq);
    for (int v = v_lo; v < v_hi; ++v) {
        const uint4 kv = *reinterpret_cast<const uint4*>(k16_base + 8 * v);
        const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t u = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
            const int a = 8 * v + e - sh;
            if ((unsigned)a >= (unsigned)A) continue;
            const bool eq = u == Tk;
            if (u > Tk || (eq && eq_idx < r_eq)) {
                sort32[out] = (u << 16) | (0xffffu - (uint32_t)out);
                sel16[out] = (unsigned short)a;
                ++out;
            }
            eq_idx += eq ? 1 : 0;
        }
    }
    const int n_sel = take_k;                                              // exactly take_k anchors are selected
    for (int i = n_sel + tid; i < sort_cap; i += NT) sort32[i] = 0u;
    __syncthreads();

    // ---- 3. request the survivors' fused rows (device or pinned host memory) -> shared memory.  RPV adjacent lanes copy one
    //         row in ONE instruction, so a row is a single coalesced 64 / 128-byte request (over PCIe: one read, not four) ---
    for (int idx = tid; idx < n_sel * RPV; idx += NT) {
        const int i = idx / RPV, v = idx % RPV;
        const int a = (int)sel16[i];
        int l = 0;
#pragma unroll
        for (int q = 1; q < TSCD_MAX_LEVELS; ++q)
            if (q < an.num_levels && a >= an.level_start[q]) l = q;
        const __half* src = reinterpret_cast<const __half*>(args.reg.ptr[l]) + (int64_t)frame * args.reg.frame_stride[l] +
                            (int64_t)(a - an.level_start[l]) * (RPV * 8);
        cp_async16(rows + idx, src + 8 * v);
    }

    // ---- 4. order by objectness (descending; ties: lower anchor id == lower compaction position first).  With cand_rank the
    //         sort is skipped: candidates stay in anchor order and the key travels to K2 as the tie-break rank ----------------
    const bool unsorted = args.cand_rank != nullptr;
    if (!unsorted) block_sort_desc64_dyn<uint32_t>(sort32, n_sel, sort_cap);
    cp_async_wait_all();
    __syncthreads();

    // ---- 5. candidate records from the staged rows -----------------------------------------------------------------
    const int64_t base = (int64_t)frame * args.cand_cap;
    for (int j = tid; j < n_sel; j += NT) {
        const int i = 0xffff - (int)(sort32[j] & 0xffffu);
        const int a = (int)sel16[i];
        float conf = -INFINITY, reg4[4] = {0.f, 0.f, 0.f, 0.f}, obj = 0.f;
        int cls_id = 0;
#pragma unroll
        for (int v = 0; v < RPV; ++v) {
            const uint4 q = rows[(size_t)i * RPV + v];
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int e = v * 8 + t;                                   // element of the fused row (compile-time)
                const float x = hbits2f((w[t >> 1] >> (16 * (t & 1))) & 0xffffu);
                if (e < 4) reg4[e] = x;
                else if (e == 4) obj = x;
                else if (e - 5 < C && x > conf) { conf = x; cls_id = e - 5; }   // first maximum wins (torch.max)
            }
        }
        if (sig) { conf = sigmoidf_ref(conf); obj = sigmoidf_ref(obj); }
        const AnchorPos p = anchor_pos(an, a);
        const float4 box = box_from_reg(reg4[0], reg4[1], reg4[2], reg4[3], p, args.apply_decode != 0);
        args.cand_idx[base + j] = a;
        reinterpret_cast<float4*>(args.cand_box)[base + j] = box;
        args.cand_score[base + j] = __fmul_rn(obj, conf);                  // post_process.py:512  obj * class_conf
        args.cand_cls[base + j] = cls_id;
        if (unsorted) args.cand_rank[base + j] = sort32[j];
    }
    if (tid == 0) args.cand_count[frame] = n_sel;
}

// Head pack: per-level conv outputs (any strided layout) -> fused rows + dense objectness plane.  One thread per anchor.
__global__ void __launch_bounds__(256) pack_head_kernel(const tscd_pack_head_args args) {
    const tscd_anchors& an = args.anchors;
    const int A = an.level_start[an.num_levels];
    const int a = blockIdx.x * blockDim.x + threadIdx.x, frame = blockIdx.y;
    if (a >= A) return;
    int l = 0;
#pragma unroll
    for (int q = 1; q < TSCD_MAX_LEVELS; ++q)
        if (q < an.num_levels && a >= an.level_start[q]) l = q;
    const int local = a - an.level_start[l];
    const __half* rp = view_ptr<__half>(args.reg, l, frame, local);
    const __half* op = view_ptr<__half>(args.obj, l, frame, local);
    const __half* cp = view_ptr<__half>(args.cls, l, frame, local);
    const int64_t rcs = args.reg.chan_stride[l], ccs = args.cls.chan_stride[l];
    const int C = args.num_classes, RP = args.row_pitch;
    __half* dst = reinterpret_cast<__half*>(args.rows) + ((int64_t)frame * A + a) * RP;
    const __half o = __ldg(op);
    for (int e0 = 0; e0 < RP; e0 += 8) {
        uint4 out;
        __half* h = reinterpret_cast<__half*>(&out);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int e = e0 + k;
            h[k] = e < 4 ? __ldg(rp + e * rcs) : (e == 4 ? o : (e < 5 + C ? __ldg(cp + (e - 5) * ccs) : __float2half_rn(0.f)));
        }
        *reinterpret_cast<uint4*>(dst + e0) = out;
    }
    reinterpret_cast<__half*>(args.obj_plane)[(int64_t)frame * args.obj_pitch + a] = o;
}

// The fused layout as the views describe it: class-contiguous rows of `rp` elements holding reg at +0 and cls at +5, one
// dense objectness plane over all levels.  Returns the row pitch in elements (32 / 64) or 0.
int fused_rows_pitch(const tscd_anchors& an, const tscd_view& reg, const tscd_view& obj, const tscd_view& cls, int num_classes,
                     int head_dtype, bool need_flat_obj) {
    if (head_dtype != TSCD_F16) return 0;
    const int64_t rp = reg.anchor_stride[0];
    if (rp != 32 && rp != 64) return 0;
    if (5 + num_classes > rp) return 0;
    for (int l = 0; l < an.num_levels; ++l) {
        const int nl = an.level_start[l + 1] - an.level_start[l];
        if (reg.anchor_stride[l] != rp || cls.anchor_stride[l] != rp || reg.chan_stride[l] != 1 || cls.chan_stride[l] != 1) return 0;
        if (reinterpret_cast<const __half*>(cls.ptr[l]) != reinterpret_cast<const __half*>(reg.ptr[l]) + 5) return 0;
        if ((reinterpret_cast<uintptr_t>(reg.ptr[l]) % (rp * 2)) || (reg.frame_stride[l] % rp) || reg.frame_stride[l] != cls.frame_stride[l]) return 0;
        if (need_flat_obj) {
            if (obj.anchor_stride[l] != 1 || obj.frame_stride[l] != obj.frame_stride[0]) return 0;
            if (l + 1 < an.num_levels && reinterpret_cast<const __half*>(obj.ptr[l + 1]) != reinterpret_cast<const __half*>(obj.ptr[l]) + nl) return 0;
        }
    }
    return (int)rp;
}

// Returns 1 when the launch was taken by the fused-row kernel, 0 when the layout / mode does not qualify, < 0 on error.
int select_rows_try(const tscd_select_args* a, cudaStream_t st) {
    if (a->mode != 0) return 0;
    const int A = a->anchors.level_start[a->anchors.num_levels];
    if (A > 65535) return 0;
    const int rp = fused_rows_pitch(a->anchors, a->reg, a->obj, a->cls, a->num_classes, a->head_dtype, true);
    if (!rp) return 0;
    const int take_cap = a->pre_k < A ? a->pre_k : A;
    if (take_cap > a->cand_cap) return 0;
    int sort_cap = kRowsThreads;
    while (sort_cap < take_cap) sort_cap <<= 1;
    if (sort_cap > 16 * kRowsThreads) return 0;
    size_t rows_bytes = (size_t)take_cap * rp * 2;
    if (rows_bytes < (size_t)(kRowsThreads / 32) * 512) rows_bytes = (size_t)(kRowsThreads / 32) * 512;   // per-warp histograms alias the rows
    const size_t smem = rows_bytes + (size_t)sort_cap * 4 + (size_t)((A + 15) & ~7) * 2 + (size_t)take_cap * 2 + 16;
    if (smem > 200 * 1024) return 0;
    cudaError_t e;
    if (rp == 32) {
        e = cudaFuncSetAttribute(select_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return TSCD_ERR_CUDA;
        select_rows_kernel<4><<<a->num_frames, kRowsThreads, smem, st>>>(*a, sort_cap, take_cap);
    } else {
        e = cudaFuncSetAttribute(select_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return TSCD_ERR_CUDA;
        select_rows_kernel<8><<<a->num_frames, kRowsThreads, smem, st>>>(*a, sort_cap, take_cap);
    }
    TSCD_CUDA_CHECK_LAUNCH();
    return 1;
}

}  // namespace tscd

extern "C" int tscd_pack_head(const tscd_pack_head_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->num_frames < 0 || a->num_classes <= 0 || !a->rows || !a->obj_plane) return TSCD_ERR_INVALID_ARG;
    if (a->head_dtype != TSCD_F16) return TSCD_ERR_UNSUPPORTED;
    if ((a->row_pitch != 32 && a->row_pitch != 64) || 5 + a->num_classes > a->row_pitch) return TSCD_ERR_INVALID_ARG;
    const int A = a->anchors.level_start[a->anchors.num_levels];
    if (A <= 0 || a->obj_pitch < A || (reinterpret_cast<uintptr_t>(a->rows) & 15)) return TSCD_ERR_INVALID_ARG;
    if (a->num_frames == 0) return TSCD_OK;
    pack_head_kernel<<<dim3((A + 255) / 256, a->num_frames), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}

// Test hook: canonical selection key and fp32 score of n fp16 bit patterns (device pointers).  Proves, exhaustively over all
// 65536 inputs, that ordering by the 16-bit key is ordering by the fp32 score (tests/test_gpu_selection.py).
extern "C" int tscd_debug_select_keys(const unsigned short* half_bits, int n, int apply_sigmoid, unsigned short* key, float* score,
                                      void* stream) {
    using namespace tscd;
    if (!half_bits || !key || !score || n < 0) return TSCD_ERR_INVALID_ARG;
    if (n == 0) return TSCD_OK;
    canon_key_debug_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(half_bits, n, apply_sigmoid, key, score);
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}
