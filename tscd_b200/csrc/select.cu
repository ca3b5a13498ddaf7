// K1: fused decode + score + per-frame proposal selection (modes A and B).
//
// Reference semantics: TSCDHead.postprocess_widx (yolox/models/tscd_head.py:1546-1693) and
// postpro_woclass (yolox/models/post_process.py:464-521); see include/tscd_b200.h.
//
// One CTA per frame.  Pass 1 streams the frame's objectness / class planes once from HBM and keeps one
// 32-bit order-preserving key per anchor in shared memory (A*4 B; 27 KB at A=6804).  Pass 2 is a block
// radix select (4 x 8-bit digits) for the k-th largest key with the documented tie-break (equal keys:
// lower anchor id first), a stable compaction, and for mode A a bitonic sort of the P survivors.  Only the
// survivors' regression / class rows are touched again (L2 hits).
#include "common.cuh"

namespace tscd {

constexpr int kSelThreads = 512;

// largest power-of-two vector width (in elements, <= 16 bytes) that the class planes of level l allow
template <typename T>
__host__ __device__ inline int classmax_vw(const tscd_select_args& a, int l) {
    const int nl = a.anchors.level_start[l + 1] - a.anchors.level_start[l];
    int vw = 16 / (int)sizeof(T);
    while (vw > 1 && ((a.cls.chan_stride[l] % vw) || (a.cls.frame_stride[l] % vw) || (nl % vw) ||
                      (reinterpret_cast<uintptr_t>(a.cls.ptr[l]) % (vw * sizeof(T)))))
        vw >>= 1;
    return vw;
}

template <typename T>
__global__ void __launch_bounds__(256) classmax_kernel(const tscd_select_args args) {
    constexpr int VMAX = 16 / sizeof(T);
    const int frame = blockIdx.y;
    const tscd_anchors& an = args.anchors;
    const int C = args.num_classes;
    static_assert(TSCD_MAX_LEVELS == 3, "level lookup below is written for three FPN levels");
    const int vw0 = classmax_vw<T>(args, 0);
    const int vw1 = an.num_levels > 1 ? classmax_vw<T>(args, 1) : 1;
    const int vw2 = an.num_levels > 2 ? classmax_vw<T>(args, 2) : 1;
    const int g1 = (an.level_start[1] - an.level_start[0]) / vw0;
    const int g2 = g1 + (an.num_levels > 1 ? (an.level_start[2] - an.level_start[1]) / vw1 : 0);
    const int g3 = g2 + (an.num_levels > 2 ? (an.level_start[3] - an.level_start[2]) / vw2 : 0);
    T* wconf = reinterpret_cast<T*>(args.ws_conf) + (int64_t)frame * args.ws_pitch;
    unsigned char* wcls = args.ws_cls + (int64_t)frame * args.ws_pitch;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g3) return;
    const int l = (g >= g1 ? 1 : 0) + (g >= g2 ? 1 : 0);
    const int vw = l == 0 ? vw0 : (l == 1 ? vw1 : vw2);
    const int a_lo = l == 0 ? an.level_start[0] : (l == 1 ? an.level_start[1] : an.level_start[2]);
    const int a0 = (g - (l == 0 ? 0 : (l == 1 ? g1 : g2))) * vw;
    const T* cp = reinterpret_cast<const T*>(args.cls.ptr[l]) + (int64_t)frame * args.cls.frame_stride[l] + a0;
    const int64_t ccs = args.cls.chan_stride[l];
    float best[VMAX];
    int bi[VMAX];
#pragma unroll
    for (int v = 0; v < VMAX; ++v) { best[v] = -INFINITY; bi[v] = 0; }
    for (int c0 = 0; c0 < C; c0 += 8) {
        uint4 raw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (c0 + u < C) {
                const T* src = cp + (int64_t)(c0 + u) * ccs;
                const int bytes = vw * (int)sizeof(T);
                if (bytes == 16) raw[u] = __ldg(reinterpret_cast<const uint4*>(src));
                else if (bytes == 8) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(src)); raw[u] = make_uint4(t.x, t.y, 0, 0); }
                else if (bytes == 4) raw[u] = make_uint4(__ldg(reinterpret_cast<const uint32_t*>(src)), 0, 0, 0);
                else raw[u] = make_uint4((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(src)), 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (c0 + u < C) {
                const T* e = reinterpret_cast<const T*>(&raw[u]);
#pragma unroll
                for (int v = 0; v < VMAX; ++v) {
                    const float x = ldf_reg(e[v]);
                    if (v < vw && x > best[v]) { best[v] = x; bi[v] = c0 + u; }     // first maximum wins (torch.max)
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VMAX; ++v) {
        if (v < vw) {
            wconf[a_lo + a0 + v] = cvt_from_float<T>(best[v]);      // exact: the maximum is one of the T-typed inputs
            wcls[a_lo + a0 + v] = (unsigned char)bi[v];
        }
    }
}

// rp: row pitch (elements) of the fused head layout (csrc/select_rows.cu; fp16 only), 0 = generic strided views
template <typename T, bool PRE>
__global__ void __launch_bounds__(kSelThreads, PRE ? 3 : 2) select_kernel(const tscd_select_args args, int sort_cap, int rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int frame = blockIdx.x;
    const tscd_anchors& an = args.anchors;
    const int A = an.level_start[an.num_levels];
    const int C = args.num_classes;
    const bool sig = args.apply_sigmoid != 0;

    const int A4 = (A + 3) & ~3;
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);                     // [A] order-preserving selection key
    float* conf_s = reinterpret_cast<float*>(keys + A4);                        // [A] class max (mode A: pre-sigmoid); absent if PRE
    const int S4 = (min(A, args.cand_cap) + 3) & ~3;
    int* sel = reinterpret_cast<int*>(conf_s + (PRE ? 0 : A4));                 // [min(A, cand_cap)] selected anchor ids
    unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(sel + S4 + (S4 & 1 ? 1 : 0));  // [pow2(pre_k)], 8-byte aligned
    unsigned char* cls_s = reinterpret_cast<unsigned char*>(sortbuf + sort_cap); // [A] class arg-max (absent if PRE)
    // PRE + fp16 logits (mode A): 16-bit order-preserving keys of the raw objectness logits.  sigmoid is monotone, so the
    // k-th largest logit locates the k-th largest score with a 2-pass radix select instead of 4 passes over fp32 keys.
    constexpr bool K16 = PRE && sizeof(T) == 2;
    unsigned short* k16 = reinterpret_cast<unsigned short*>(sortbuf + sort_cap);   // [A] (aliases cls_s, which PRE does not use)
    __shared__ SelSmem s;

    // ---- pass 1: stream objectness + class planes once; key / class max / arg-max per anchor ----------
    // Planar layouts (anchor stride 1: NCHW conv outputs) are read with 16-byte vector loads, V anchors per
    // thread, all C+1 planes of the group in flight; other layouts take the per-anchor path.
    int n_ge = 0;  // mode B: anchors with score >= conf_thresh
    const bool modeB = args.mode == 1;
    constexpr int VMAX = 16 / sizeof(T);
    // vector groups of all levels share one index space, so the small levels do not cost extra latency rounds
    int nvec_l[TSCD_MAX_LEVELS], goff[TSCD_MAX_LEVELS + 1];
    goff[0] = 0;
#pragma unroll
    for (int l = 0; l < TSCD_MAX_LEVELS; ++l) {
        nvec_l[l] = 0;
        if (l < an.num_levels) {
            const int nl = an.level_start[l + 1] - an.level_start[l];
            const T* op = reinterpret_cast<const T*>(args.obj.ptr[l]) + (int64_t)frame * args.obj.frame_stride[l];
            const T* cp = reinterpret_cast<const T*>(args.cls.ptr[l]) + (int64_t)frame * args.cls.frame_stride[l];
            const bool planar = args.obj.anchor_stride[l] == 1 && args.cls.anchor_stride[l] == 1 &&
                                ((reinterpret_cast<uintptr_t>(op) | reinterpret_cast<uintptr_t>(cp)) & 15) == 0 &&
                                (args.cls.chan_stride[l] % VMAX) == 0 && (args.obj.frame_stride[l] % VMAX) == 0 &&
                                (args.cls.frame_stride[l] % VMAX) == 0;
            nvec_l[l] = (planar && !(rp && !PRE)) ? nl / VMAX : 0;
        }
        goff[l + 1] = goff[l] + nvec_l[l];
    }
    for (int g = threadIdx.x; g < goff[TSCD_MAX_LEVELS]; g += blockDim.x) {
        static_assert(TSCD_MAX_LEVELS == 3, "level lookup below is written for three FPN levels");
        const int l = (g >= goff[1] ? 1 : 0) + (g >= goff[2] ? 1 : 0);
        const int a_lo = l == 0 ? an.level_start[0] : (l == 1 ? an.level_start[1] : an.level_start[2]);
        const int a0 = (g - (l == 0 ? 0 : (l == 1 ? goff[1] : goff[2]))) * VMAX;
        const T* op = reinterpret_cast<const T*>(args.obj.ptr[l]) + (int64_t)frame * args.obj.frame_stride[l] + a0;
        const T* cp = reinterpret_cast<const T*>(args.cls.ptr[l]) + (int64_t)frame * args.cls.frame_stride[l] + a0;
        const int64_t ccs = args.cls.chan_stride[l];
        float ob[VMAX], best[VMAX];
        int bi[VMAX];
        const uint4 oraw = __ldg(reinterpret_cast<const uint4*>(op));
#pragma unroll
        for (int v = 0; v < VMAX; ++v) { best[v] = -INFINITY; bi[v] = 0; }
        // class planes in batches of 8 independent 16-byte loads (128 B in flight per thread)
        for (int c0 = 0; c0 < (PRE ? 0 : C); c0 += 8) {
            uint4 raw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < C) raw[u] = __ldg(reinterpret_cast<const uint4*>(cp + (int64_t)(c0 + u) * ccs));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (c0 + u < C) {
                    const T* e = reinterpret_cast<const T*>(&raw[u]);
#pragma unroll
                    for (int v = 0; v < VMAX; ++v) {
                        const float x = ldf_reg(e[v]);
                        if (x > best[v]) { best[v] = x; bi[v] = c0 + u; }     // first maximum wins (torch.max)
                    }
                }
            }
        }
        {
            const T* e = reinterpret_cast<const T*>(&oraw);
#pragma unroll
            for (int v = 0; v < VMAX; ++v) ob[v] = ldf_reg(e[v]);
        }
#pragma unroll
        for (int v = 0; v < VMAX; ++v) {
            const int a = a_lo + a0 + v;
            float o = ob[v], b = best[v];
            if (sig) o = sigmoidf_ref(o);
            float key = o;
            if (modeB) {
                if (sig) b = sigmoidf_ref(b);
                key = __fmul_rn(o, b);                            // tscd_head.py:1591  obj * class_conf
                n_ge += (key >= args.conf_thresh) ? 1 : 0;
            }
            keys[a] = f2ord(key);
            if (!PRE) { conf_s[a] = b; cls_s[a] = (unsigned char)bi[v]; }
            if (K16) {
                const unsigned short hb = reinterpret_cast<const unsigned short*>(&oraw)[v];
                k16[a] = (hb & 0x8000u) ? (unsigned short)~hb : (unsigned short)(hb | 0x8000u);
            }
        }
    }
    // fused 64 / 128-byte rows [reg4|obj|cls C|pad] (mode B streams every anchor's row): one thread per anchor, the row as
    // 16-byte vector loads -- consecutive lanes read consecutive rows, so the four loads of a warp cover whole sectors
    if (rp && sizeof(T) == 2 && !PRE) {
        for (int l = 0; l < an.num_levels; ++l) {
            const int a_lo = an.level_start[l], nl = an.level_start[l + 1] - a_lo;
            const uint4* rowp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(args.reg.ptr[l]) + (int64_t)frame * args.reg.frame_stride[l]);
            const int rpv = rp / 8;
            for (int i = threadIdx.x; i < nl; i += blockDim.x) {
                float b = -INFINITY, o = 0.f;
                int bidx = 0;
                for (int v = 0; v < rpv; ++v) {
                    const uint4 q = __ldg(rowp + (int64_t)i * rpv + v);
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const int e = v * 8 + t;
                        const float x = __half2float(__ushort_as_half((unsigned short)((w[t >> 1] >> (16 * (t & 1))) & 0xffffu)));
                        if (e == 4) o = x;
                        else if (e >= 5 && e - 5 < C && x > b) { b = x; bidx = e - 5; }   // first maximum wins (torch.max)
                    }
                }
                if (sig) o = sigmoidf_ref(o);
                float key = o;
                if (modeB) {
                    if (sig) b = sigmoidf_ref(b);
                    key = __fmul_rn(o, b);
                    n_ge += (key >= args.conf_thresh) ? 1 : 0;
                }
                const int a = a_lo + i;
                keys[a] = f2ord(key);
                conf_s[a] = b;
                cls_s[a] = (unsigned char)bidx;
            }
        }
    }
    for (int l = 0; l < (rp && sizeof(T) == 2 && !PRE ? 0 : an.num_levels); ++l) {
        const int a_lo = an.level_start[l], nl = an.level_start[l + 1] - a_lo;
        const T* op = reinterpret_cast<const T*>(args.obj.ptr[l]) + (int64_t)frame * args.obj.frame_stride[l];
        const T* cp = reinterpret_cast<const T*>(args.cls.ptr[l]) + (int64_t)frame * args.cls.frame_stride[l];
        const int64_t ccs = args.cls.chan_stride[l];
        const int nvec = nvec_l[l];
        for (int i = nvec * VMAX + threadIdx.x; i < nl; i += blockDim.x) {      // tail / non-planar layouts
            const int a = a_lo + i;
            const T oraw1 = __ldg(op + (int64_t)i * args.obj.anchor_stride[l]);
            float o = ldf_reg(oraw1);
            if (K16) {
                const unsigned short hb = *reinterpret_cast<const unsigned short*>(&oraw1);
                k16[a] = (hb & 0x8000u) ? (unsigned short)~hb : (unsigned short)(hb | 0x8000u);
            }
            const T* c = cp + (int64_t)i * args.cls.anchor_stride[l];
            float b = ldf(c);
            int bidx = 0;
            for (int k = 1; k < (PRE ? 0 : C); ++k) {
                const float x = ldf(c + k * ccs);
                if (x > b) { b = x; bidx = k; }
            }
            if (sig) o = sigmoidf_ref(o);
            float key = o;
            if (modeB) {
                if (sig) b = sigmoidf_ref(b);
                key = __fmul_rn(o, b);
                n_ge += (key >= args.conf_thresh) ? 1 : 0;
            }
            keys[a] = f2ord(key);
            if (!PRE) { conf_s[a] = b; cls_s[a] = (unsigned char)bidx; }
        }
    }
    __syncthreads();

    // ---- how many to take --------------------------------------------------------------------------
    int take_k = 0;  // 0 = plain threshold mask (mode B, no limit triggered)
    if (args.mode == 0) {
        take_k = min(args.pre_k, A);
    } else {
        int tot;
        block_excl_scan(n_ge, s.scan, &tot);
        int cnt = tot;
        if (args.minimal_limit != 0 && cnt < args.minimal_limit) {  // :1594-1599 (top-min is a superset of the mask)
            take_k = min(args.minimal_limit, A);
            cnt = take_k;
        }
        if (args.maximal_limit != 0 && cnt > args.maximal_limit) {  // :1600-1607
            take_k = min(args.maximal_limit, A);
        }
    }

    uint32_t Tk = f2ord(args.conf_thresh);
    int r_eq = 0x7fffffff;  // threshold mode: every key == Tk is taken
    bool req_from_count = false;
    if (take_k > 0) {
        if (K16 && args.mode == 0) {
            uint32_t x_k;
            int dummy;
            radix_select_kth<unsigned short>(k16, A, take_k, &s, &x_k, &dummy);
            // score key of the k-th logit (every anchor with this logit has the same score); equal SCORES of different
            // logits (sigmoid saturation) are resolved by the fp32 keys below: r_eq = k - #(score > T)
            for (int a = threadIdx.x; a < A; a += blockDim.x)
                if ((uint32_t)k16[a] == x_k) s.misc[2] = (int)keys[a];
            __syncthreads();
            Tk = (uint32_t)s.misc[2];
            req_from_count = true;
        } else {
            radix_select_kth<uint32_t>(keys, A, take_k, &s, &Tk, &r_eq);
        }
    }

    // ---- stable compaction in ascending anchor order ---------------------------------------------------
    const int chunk = (A + blockDim.x - 1) / blockDim.x;
    const int lo = min(threadIdx.x * chunk, A), hi = min(lo + chunk, A);
    int n_gt = 0, n_eq = 0;
    for (int a = lo; a < hi; ++a) {
        uint32_t u = keys[a];
        n_gt += (u > Tk);
        n_eq += (u == Tk);
    }
    int tot_eq, tot_gt;
    if (req_from_count) {
        int tot_strict;
        block_excl_scan(n_gt, s.scan, &tot_strict);
        r_eq = take_k - tot_strict;
    }
    int eq_before = block_excl_scan(n_eq, s.scan, &tot_eq);
    int take_eq = max(0, min(n_eq, r_eq - eq_before));  // lowest anchor ids among the ties
    int out_before = block_excl_scan(n_gt + take_eq, s.scan, &tot_gt);
    const int n_sel_raw = tot_gt;
    {
        int o = out_before, e = eq_before;
        for (int a = lo; a < hi; ++a) {
            uint32_t u = keys[a];
            bool t = (u > Tk);
            if (u == Tk) { t = (e < r_eq); ++e; }
            if (t) { if (o < args.cand_cap) sel[o] = a; ++o; }
        }
    }
    __syncthreads();
    const int n_sel = min(n_sel_raw, args.cand_cap);
    if (n_sel_raw > args.cand_cap && args.status && threadIdx.x == 0) atomicMin(args.status, TSCD_ERR_CAPACITY);

    // ---- mode A: order by objectness, descending, ties lower anchor id first ---------------------------
    if (args.mode == 0) {
        bool sorted = false;
        if (K16 && A <= 65536) {
            // 32-bit keys (logit order | inverted anchor id): half the compare-exchange work of the 64-bit sort.  Valid
            // unless two selected anchors with DIFFERENT logits share a score (sigmoid saturation): checked on the
            // sorted list (such a pair is adjacent because sigmoid is monotone); then the 64-bit sort redoes the order.
            uint32_t* sort32 = reinterpret_cast<uint32_t*>(sortbuf);
            for (int i = threadIdx.x; i < n_sel; i += blockDim.x) {
                const int a = sel[i];
                sort32[i] = ((uint32_t)k16[a] << 16) | (0xffffu - (uint32_t)a);
            }
            __syncthreads();
            block_sort_desc64_dyn<uint32_t>(sort32, n_sel, sort_cap);
            int bad = 0;
            for (int i = threadIdx.x; i + 1 < n_sel; i += blockDim.x) {
                const int a0 = 0xffff - (int)(sort32[i] & 0xffffu), a1 = 0xffff - (int)(sort32[i + 1] & 0xffffu);
                bad |= (k16[a0] != k16[a1] && keys[a0] == keys[a1]) ? 1 : 0;
            }
            if (!__syncthreads_or(bad)) {
                for (int i = threadIdx.x; i < n_sel; i += blockDim.x) sel[i] = 0xffff - (int)(sort32[i] & 0xffffu);
                sorted = true;
            }
            __syncthreads();
        }
        if (!sorted) {
            for (int i = threadIdx.x; i < n_sel; i += blockDim.x) {
                const int a = sel[i];
                sortbuf[i] = ((unsigned long long)keys[a] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)a);
            }
            __syncthreads();
            block_sort_desc64_dyn<unsigned long long>(sortbuf, n_sel, sort_cap);      // registers + shuffles; smem only for distances >= 32 * E
            for (int i = threadIdx.x; i < n_sel; i += blockDim.x)
                sel[i] = (int)(0xffffffffu - (uint32_t)(sortbuf[i] & 0xffffffffull));
        }
        __syncthreads();
    }

    // ---- candidate records: only the regression rows of the survivors are touched again ------------------
    const int64_t base = (int64_t)frame * args.cand_cap;
    for (int i = threadIdx.x; i < n_sel; i += blockDim.x) {
        const int a = sel[i];
        AnchorPos p = anchor_pos(an, a);
        float score;
        int cls_id = 0;
        if (modeB) {
            score = ord2f(keys[a]);
            cls_id = (int)cls_s[a];
        } else {
            float conf;
            if (!PRE) {
                conf = conf_s[a];
                cls_id = (int)cls_s[a];
            } else if (args.ws_conf) {       // class max of every anchor was streamed by classmax_kernel
                conf = ldf(reinterpret_cast<const T*>(args.ws_conf) + (int64_t)frame * args.ws_pitch + a);
                cls_id = (int)args.ws_cls[(int64_t)frame * args.ws_pitch + a];
            } else {                         // class-contiguous (channels_last) logits: max over the survivor's own row only
                const T* c = view_ptr<T>(args.cls, p.level, frame, p.local);
                const int64_t ccs = args.cls.chan_stride[p.level];
                conf = -INFINITY;
                for (int k0 = 0; k0 < C; k0 += 8) {                   // 8 independent loads in flight
                    float x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = k0 + u < C ? ldf(c + (k0 + u) * ccs) : -INFINITY;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (x[u] > conf) { conf = x[u]; cls_id = k0 + u; }   // first maximum wins (torch.max)
                }
            }
            if (sig) conf = sigmoidf_ref(conf);
            score = __fmul_rn(ord2f(keys[a]), conf);                  // post_process.py:512  obj * class_conf
        }
        float4 box = anchor_box<T>(args.reg, p, frame, args.apply_decode != 0);
        args.cand_idx[base + i] = a;
        reinterpret_cast<float4*>(args.cand_box)[base + i] = box;
        args.cand_score[base + i] = score;
        args.cand_cls[base + i] = cls_id;
    }
    if (threadIdx.x == 0) args.cand_count[frame] = n_sel;
}

}  // namespace tscd

extern "C" int tscd_select(const tscd_select_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->num_frames < 0 || a->num_classes <= 0 || a->cand_cap <= 0) return TSCD_ERR_INVALID_ARG;
    if (a->anchors.num_levels < 1 || a->anchors.num_levels > TSCD_MAX_LEVELS) return TSCD_ERR_INVALID_ARG;
    if (a->mode != 0 && a->mode != 1) return TSCD_ERR_INVALID_ARG;
    if (a->mode == 0 && a->pre_k <= 0) return TSCD_ERR_INVALID_ARG;
    if (a->num_frames == 0) return TSCD_OK;
    const int A = a->anchors.level_start[a->anchors.num_levels];
    if (A <= 0) return TSCD_ERR_INVALID_ARG;
    int n64 = 1;
    if (a->mode == 0) while (n64 < (a->pre_k < A ? a->pre_k : A)) n64 <<= 1;
    if (n64 < kSelThreads) n64 = kSelThreads;            // the block sort works on E * blockDim.x keys
    if (n64 > 16 * kSelThreads) return TSCD_ERR_CAPACITY;
    if (a->num_classes > 255) return TSCD_ERR_UNSUPPORTED;
    const int selcap = ((A < a->cand_cap ? A : a->cand_cap) + 3) & ~3;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {   // fused 64 / 128-byte rows + dense objectness plane (the drop-in head's layout): csrc/select_rows.cu
        const int rc = select_rows_try(a, st);
        if (rc != 0) return rc < 0 ? rc : TSCD_OK;
        if (a->cand_rank) return TSCD_ERR_UNSUPPORTED;       // the unsorted / rank output exists for the fused-row kernel only
    }
    // mode A with a workspace and planar (anchor-contiguous) class planes: class max as a separate streaming kernel
    // mode A needs the class max of the ~pre_k survivors only.  Class-contiguous logits (channels_last conv outputs: an
    // anchor's C logits are one 2C-byte row) are therefore read for the survivors alone -- 89 % of the class logits never
    // leave HBM; anchor-contiguous (NCHW) planes cannot be gathered sector-efficiently and are streamed by classmax_kernel.
    bool late = a->mode == 0;
    for (int l = 0; late && l < a->anchors.num_levels; ++l) late = a->cls.chan_stride[l] == 1;
    bool pre = !late && a->mode == 0 && a->ws_conf && a->ws_cls && a->ws_pitch >= A && (a->head_dtype == TSCD_F16 || a->head_dtype == TSCD_F32);
    const int vmax = a->head_dtype == TSCD_F16 ? 8 : 4, esz = a->head_dtype == TSCD_F16 ? 2 : 4;
    int groups = 0;
    for (int l = 0; pre && l < a->anchors.num_levels; ++l) {
        const int nl = a->anchors.level_start[l + 1] - a->anchors.level_start[l];
        pre = a->cls.anchor_stride[l] == 1;                      // anchor-contiguous planes (NCHW conv outputs)
        groups += nl / (a->head_dtype == TSCD_F16 ? classmax_vw<__half>(*a, l) : classmax_vw<float>(*a, l));
    }
    (void)vmax;
    (void)esz;
    const bool light = pre || late;          // select_kernel<T, true>: no class data in shared memory
    const size_t smem = (size_t)((A + 3) & ~3) * (light ? 4 : 8) + (size_t)selcap * 4 + 8 + (size_t)n64 * 8 + (light ? (size_t)2 * ((A + 15) & ~15) : (size_t)((A + 15) & ~15));
    if (smem > 200 * 1024) return TSCD_ERR_CAPACITY;
    cudaError_t e;
    if (pre) {
        const dim3 grid((groups + 255) / 256 > 0 ? (groups + 255) / 256 : 1, a->num_frames);
        if (a->head_dtype == TSCD_F32) classmax_kernel<float><<<grid, 256, 0, st>>>(*a);
        else classmax_kernel<__half><<<grid, 256, 0, st>>>(*a);
        TSCD_CUDA_CHECK_LAUNCH();
    }
#define TSCD_LAUNCH_SELECT(TT, PP)                                                                                   \
    do {                                                                                                             \
        e = cudaFuncSetAttribute(select_kernel<TT, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return TSCD_ERR_CUDA;                                                                  \
        select_kernel<TT, PP><<<a->num_frames, kSelThreads, smem, st>>>(*a, n64, rows_rp);                           \
    } while (0)
    const int rows_rp = fused_rows_pitch(a->anchors, a->reg, a->obj, a->cls, a->num_classes, a->head_dtype, false);
    tscd_select_args la = *a;                 // late path: no workspace reads
    if (late) { la.ws_conf = nullptr; la.ws_cls = nullptr; }
    a = &la;
    if (a->head_dtype == TSCD_F32) {
        if (light) TSCD_LAUNCH_SELECT(float, true); else TSCD_LAUNCH_SELECT(float, false);
    } else if (a->head_dtype == TSCD_F16) {
        if (light) TSCD_LAUNCH_SELECT(__half, true); else TSCD_LAUNCH_SELECT(__half, false);
    } else {
        return TSCD_ERR_UNSUPPORTED;
    }
#undef TSCD_LAUNCH_SELECT
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}
