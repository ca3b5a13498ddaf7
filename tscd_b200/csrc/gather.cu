// K3: per-frame counts -> bank offsets, reference rows for the kept proposals, and the gather of their
// classification / regression / edge feature rows into the packed clip bank.
//
// Reference: row build tscd_head.py:1581-1582 (+1670-1684), find_feature_score tscd_head.py:976-1006.
// One warp per kept proposal: lanes read the class scores (warp arg-max, first maximum wins like
// torch.max), lane 0 decodes the box, then the warp copies the three D-channel feature rows with 16-byte
// vector loads when the plane is channel-contiguous (channels_last conv outputs), falling back to strided
// element loads for NCHW planes.
#include <type_traits>

#include "common.cuh"
#include "tc.cuh"

namespace tscd {

// 1-D bulk copies of the TMA engine (cp.async.bulk): global / pinned host -> shared with mbarrier completion, shared -> global.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(tc::smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(tc::smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

__global__ void count_scan_kernel(int num_frames, int use_keep, int max_keep, const int32_t* cand_count,
                                  const int32_t* keep_count, int32_t* sel_count, int32_t* row_off) {
    __shared__ int scratch[40];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < num_frames; base += blockDim.x) {
        int f = base + threadIdx.x;
        int c = 0;
        if (f < num_frames) {
            c = use_keep ? keep_count[f] : cand_count[f];
            c = min(c, max_keep);
            sel_count[f] = c;
        }
        int tot;
        int ex = block_excl_scan(c, scratch, &tot);
        int carry = carry_s;
        if (f < num_frames) row_off[f] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_off[num_frames] = carry_s;
}

__device__ __forceinline__ float ld_any(const void* base, int dtype, int64_t idx) {
    if (dtype == TSCD_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
    if (dtype == TSCD_F16) return __half2float(__ldg(reinterpret_cast<const __half*>(base) + idx));
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
}

template <typename TF, typename TB>
__device__ __forceinline__ void copy_feature_row(const tscd_view& v, int level, int frame, int local, int D, TB* dst,
                                                 int lane) {
    const TF* src = reinterpret_cast<const TF*>(v.ptr[level]) + (int64_t)frame * v.frame_stride[level] +
                    (int64_t)local * v.anchor_stride[level];
    const int64_t cs = v.chan_stride[level];
    constexpr int VEC = 16 / sizeof(TF);
    if (cs == 1 && (D % VEC) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        for (int c0 = lane * VEC; c0 < D; c0 += 32 * VEC) {
            uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + c0));
            const TF* e = reinterpret_cast<const TF*>(&raw);
            if (sizeof(TF) == sizeof(TB) && ((reinterpret_cast<uintptr_t>(dst + c0) & 15) == 0)) {
                // same width: convert in registers, one 16-byte store
                uint4 out;
                TB* o = reinterpret_cast<TB*>(&out);
#pragma unroll
                for (int k = 0; k < VEC; ++k) o[k] = cvt_from_float<TB>(ldf_reg(e[k]));
                *reinterpret_cast<uint4*>(dst + c0) = out;
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) dst[c0 + k] = cvt_from_float<TB>(ldf_reg(e[k]));
            }
        }
    } else {
        for (int c = lane; c < D; c += 32) dst[c] = cvt_from_float<TB>(ldf(src + c * cs));
    }
}

// rp: row pitch (elements) of the fused head layout (csrc/select_rows.cu), 0 = generic strided views
// bulk: TSCD-L fast layout (256 channel-contiguous 16-bit features, bank of the same type): the three feature rows of every
//       kept proposal are staged through shared memory by the TMA engine -- one lane issues three 512-byte cp.async.bulk loads
//       per row (device memory or, for forward_host, pinned host memory), and since the kept rows of a frame are CONSECUTIVE in
//       the packed bank, each plane leaves as ONE bulk store of n x 512 bytes; the warps meanwhile build the reference rows
//       (class scores, box) of the same proposals.  No feature byte passes through a register.
constexpr int kGatherBulkRows = 32;          // rows staged per batch: 3 planes x 32 x 512 B = 48 KB (16-row batches measured slower)
constexpr int kGatherMaxRows = 512;          // kept rows per frame whose anchor ids are staged in shared memory

template <typename TF, typename TB>
__global__ void __launch_bounds__(256) rows_gather_kernel(const tscd_gather_args args, int rp, int bulk) {
    extern __shared__ __align__(128) unsigned char gsm[];
    __shared__ __align__(8) uint64_t gbar;
    const int frame = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n = args.sel_count[frame];
    const int row0 = args.row_off[frame];
    // kept position -> anchor id of every row of the frame, resolved once by the whole CTA (two dependent loads) instead of by every
    // warp in front of its own row loads
    __shared__ int s_anchor[kGatherMaxRows];
    const bool staged = gridDim.y == 1 && n <= kGatherMaxRows;
    if (staged)
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const int pos = args.use_keep ? args.keep[(int64_t)frame * args.max_keep + j] : j;
            s_anchor[j] = args.cand_idx[(int64_t)frame * args.cand_cap + pos];
        }
    if (bulk && threadIdx.x == 0) { tc::mbar_init(&gbar, 1); tc::fence_barrier_init(); }
    if (bulk || staged) __syncthreads();
    const int C = args.num_classes;
    const int W = 7 + C;
    const tscd_anchors& an = args.anchors;
    const int hd = args.head_dtype;
    // a null edge view: the edge rows are computed from the regression features afterwards (csrc/edge.cu), bank_edge is left alone
    const int np = args.feat_edge.ptr[0] ? 3 : 2;
    if (bulk && n <= kGatherBulkRows) {
        // Frames of <= 32 kept rows (mode A: 30): ONE batch whose bulk loads are issued by ALL eight warps, four rows each --
        // cp.async.bulk is a uniform-datapath instruction, so the lanes of a warp issue theirs one after the other (~100 clk
        // each); 90 loads from one warp were a third of the CTA's lifetime.  Warp 0 then waits for the bytes and stores the planes.
        TB* const banks[3] = {reinterpret_cast<TB*>(args.bank_cls), reinterpret_cast<TB*>(args.bank_reg), reinterpret_cast<TB*>(args.bank_edge)};
        if (threadIdx.x == 0) tc::mbar_expect_tx(&gbar, (uint32_t)(n * np) * 512u);
        const int j = wid * 4 + lane;
        if (lane < 4 && j < n) {
            const AnchorPos p = anchor_pos(args.anchors, s_anchor[j]);
            bulk_load(gsm + (0 * kGatherBulkRows + j) * 512, view_ptr<TF>(args.feat_cls, p.level, frame, p.local), 512, &gbar);
            bulk_load(gsm + (1 * kGatherBulkRows + j) * 512, view_ptr<TF>(args.feat_reg, p.level, frame, p.local), 512, &gbar);
            if (np == 3) bulk_load(gsm + (2 * kGatherBulkRows + j) * 512, view_ptr<TF>(args.feat_edge, p.level, frame, p.local), 512, &gbar);
        }
        if (wid == 0 && n > 0) {
            tc::mbar_wait(&gbar, 0u, 700);
            tc::fence_proxy_async();
            if (lane < np) {
                bulk_store(banks[lane] + (int64_t)row0 * 256, gsm + lane * kGatherBulkRows * 512, (uint32_t)n * 512u);
                bulk_commit();
                bulk_wait_read();
            }
            __syncwarp();
        }
    } else if (bulk && wid == 0) {
        TB* const banks[3] = {reinterpret_cast<TB*>(args.bank_cls), reinterpret_cast<TB*>(args.bank_reg), reinterpret_cast<TB*>(args.bank_edge)};
        for (int b0 = 0; b0 < n; b0 += kGatherBulkRows) {
            const int nb = min(kGatherBulkRows, n - b0);
            if (lane == 0) tc::mbar_expect_tx(&gbar, (uint32_t)(nb * np) * 512u);
            __syncwarp();
            if (lane < nb) {
                const int j = b0 + lane;
                int a;
                if (staged) a = s_anchor[j];
                else {
                    const int pos = args.use_keep ? args.keep[(int64_t)frame * args.max_keep + j] : j;
                    a = args.cand_idx[(int64_t)frame * args.cand_cap + pos];
                }
                const AnchorPos p = anchor_pos(args.anchors, a);
                bulk_load(gsm + (0 * kGatherBulkRows + lane) * 512, view_ptr<TF>(args.feat_cls, p.level, frame, p.local), 512, &gbar);
                bulk_load(gsm + (1 * kGatherBulkRows + lane) * 512, view_ptr<TF>(args.feat_reg, p.level, frame, p.local), 512, &gbar);
                if (np == 3) bulk_load(gsm + (2 * kGatherBulkRows + lane) * 512, view_ptr<TF>(args.feat_edge, p.level, frame, p.local), 512, &gbar);
            }
            tc::mbar_wait(&gbar, (uint32_t)((b0 / kGatherBulkRows) & 1), 700);
            tc::fence_proxy_async();
            if (lane < np) {         // the batch's rows are consecutive in the packed bank: one store per plane
                bulk_store(banks[lane] + (int64_t)(row0 + b0) * 256, gsm + lane * kGatherBulkRows * 512, (uint32_t)nb * 512u);
                bulk_commit();
                bulk_wait_read();    // the staging buffer may be overwritten by the next batch
            }
            __syncwarp();
        }
    }
    for (int j = wid + blockIdx.y * nw; j < n; j += nw * gridDim.y) {
        int a;
        if (staged) a = s_anchor[j];
        else {
            const int pos = args.use_keep ? args.keep[(int64_t)frame * args.max_keep + j] : j;
            a = args.cand_idx[(int64_t)frame * args.cand_cap + pos];
        }
        AnchorPos p = anchor_pos(an, a);
        float* row = args.sel_rows + ((int64_t)frame * args.max_keep + j) * W;
        const int64_t r = (int64_t)(row0 + j) * args.feat_dim;
        // fast path (TSCD-L: 256 channel-contiguous 16-bit features, same bank type): the three feature rows are requested FIRST
        // (three independent 16-byte loads per lane), so their latency overlaps the dependent class / objectness / box loads below
        bool fast = false;
        uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0, v2 = v0;
        if (!bulk && std::is_same<TF, TB>::value && sizeof(TF) == 2 && args.feat_dim == 256 &&
            args.feat_cls.chan_stride[p.level] == 1 && args.feat_reg.chan_stride[p.level] == 1 && (np == 2 || args.feat_edge.chan_stride[p.level] == 1)) {
            const TF* s0 = view_ptr<TF>(args.feat_cls, p.level, frame, p.local) + lane * 8;
            const TF* s1 = view_ptr<TF>(args.feat_reg, p.level, frame, p.local) + lane * 8;
            const TF* s2 = np == 3 ? view_ptr<TF>(args.feat_edge, p.level, frame, p.local) + lane * 8 : s1;
            if (((reinterpret_cast<uintptr_t>(s0) | reinterpret_cast<uintptr_t>(s1) | reinterpret_cast<uintptr_t>(s2)) & 15) == 0) {
                fast = true;
                v0 = __ldg(reinterpret_cast<const uint4*>(s0));
                v1 = __ldg(reinterpret_cast<const uint4*>(s1));
                if (np == 3) v2 = __ldg(reinterpret_cast<const uint4*>(s2));
            }
        }
        // class scores + arg-max (first maximum)
        float best = -INFINITY;
        int bi = 0x7fffffff;
        float obj_raw = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
        if (rp) {
            // fused layout: the anchor's whole head output is one aligned 64 / 128-byte row -> ONE coalesced load instruction
            // (a single PCIe read when the rows live in pinned host memory), elements handed around with shuffles
            const __half* rowp = reinterpret_cast<const __half*>(args.reg.ptr[p.level]) + (int64_t)frame * args.reg.frame_stride[p.level] +
                                 (int64_t)p.local * rp;
            float h0, h1 = 0.f;
            if (rp == 32) {
                h0 = __half2float(__ldg(rowp + lane));
            } else {
                const __half2 t = __ldg(reinterpret_cast<const __half2*>(rowp) + lane);
                h0 = __low2float(t);
                h1 = __high2float(t);
            }
            auto elem = [&](int e) -> float {                 // element e of the row (e may differ per lane)
                if (rp == 32) return __shfl_sync(0xffffffffu, h0, e & 31);
                const float a0 = __shfl_sync(0xffffffffu, h0, (e >> 1) & 31), a1 = __shfl_sync(0xffffffffu, h1, (e >> 1) & 31);
                return (e & 1) ? a1 : a0;
            };
            r0 = elem(0); r1 = elem(1); r2 = elem(2); r3 = elem(3);
            obj_raw = elem(4);
            for (int c0 = 0; c0 < C; c0 += 32) {              // uniform trip count: the shuffles need the whole warp
                const int c = c0 + lane;
                float v = elem(5 + (c < C ? c : 0));
                if (c < C) {
                    if (args.apply_sigmoid) v = sigmoidf_ref(v);
                    row[7 + c] = v;
                    if (v > best) { best = v; bi = c; }
                }
            }
        } else {
            const int64_t cbase = (int64_t)frame * args.cls.frame_stride[p.level] + (int64_t)p.local * args.cls.anchor_stride[p.level];
            if (lane == 0)
                obj_raw = ld_any(args.obj.ptr[p.level], hd,
                                 (int64_t)frame * args.obj.frame_stride[p.level] + (int64_t)p.local * args.obj.anchor_stride[p.level]);
            for (int c = lane; c < C; c += 32) {
                float v = ld_any(args.cls.ptr[p.level], hd, cbase + c * args.cls.chan_stride[p.level]);
                if (args.apply_sigmoid) v = sigmoidf_ref(v);
                row[7 + c] = v;
                if (v > best) { best = v; bi = c; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) {
            float obj = obj_raw;
            if (args.apply_sigmoid) obj = sigmoidf_ref(obj);
            float4 box = rp ? box_from_reg(r0, r1, r2, r3, p, args.apply_decode != 0)
                            : ((hd == TSCD_F32) ? anchor_box<float>(args.reg, p, frame, args.apply_decode != 0)
                                                : anchor_box<__half>(args.reg, p, frame, args.apply_decode != 0));
            row[0] = box.x; row[1] = box.y; row[2] = box.z; row[3] = box.w;
            row[4] = obj; row[5] = best; row[6] = (float)bi;
            args.sel_idx[(int64_t)frame * args.max_keep + j] = a;
            const int rr = row0 + j;
            args.bank_score[rr] = best;
            args.bank_fg[rr] = obj;
            reinterpret_cast<float4*>(args.bank_box)[rr] = box;
        }
        if (bulk) continue;                            // features travel through the TMA engine (below)
        if (fast) {                                    // identical 16-bit types: raw copy
            *reinterpret_cast<uint4*>(reinterpret_cast<TB*>(args.bank_cls) + r + lane * 8) = v0;
            *reinterpret_cast<uint4*>(reinterpret_cast<TB*>(args.bank_reg) + r + lane * 8) = v1;
            if (np == 3) *reinterpret_cast<uint4*>(reinterpret_cast<TB*>(args.bank_edge) + r + lane * 8) = v2;
            continue;
        }
        copy_feature_row<TF, TB>(args.feat_cls, p.level, frame, p.local, args.feat_dim, reinterpret_cast<TB*>(args.bank_cls) + r, lane);
        copy_feature_row<TF, TB>(args.feat_reg, p.level, frame, p.local, args.feat_dim, reinterpret_cast<TB*>(args.bank_reg) + r, lane);
        if (np == 3) copy_feature_row<TF, TB>(args.feat_edge, p.level, frame, p.local, args.feat_dim, reinterpret_cast<TB*>(args.bank_edge) + r, lane);
    }
}

// the TMA-bulk path needs 512-byte channel-contiguous 16-bit feature rows, 16-byte aligned, and a bank of the same type
static bool gather_bulk_ok(const tscd_gather_args* a) {
    if (a->feat_dim != 256 || a->feat_dtype != a->bank_dtype || (a->feat_dtype != TSCD_F16 && a->feat_dtype != TSCD_BF16)) return false;
    const tscd_view* vs[3] = {&a->feat_cls, &a->feat_reg, &a->feat_edge};
    const int np = a->feat_edge.ptr[0] ? 3 : 2;
    for (int v = 0; v < np; ++v)
        for (int l = 0; l < a->anchors.num_levels; ++l) {
            if (vs[v]->chan_stride[l] != 1 || (reinterpret_cast<uintptr_t>(vs[v]->ptr[l]) & 15)) return false;
            if ((vs[v]->frame_stride[l] % 8) || (vs[v]->anchor_stride[l] % 8)) return false;
        }
    const void* bs[3] = {a->bank_cls, a->bank_reg, a->bank_edge};
    for (int v = 0; v < 3; ++v)
        if (reinterpret_cast<uintptr_t>(bs[v]) & 15) return false;
    return true;
}

template <typename TF, typename TB>
static int launch_gather_kernel(const tscd_gather_args* a, dim3 grid, int rp, int bulk, cudaStream_t st) {
    const size_t sm = bulk ? (size_t)3 * kGatherBulkRows * 512 : 0;
    if (bulk) {
        if (cudaFuncSetAttribute(rows_gather_kernel<TF, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return TSCD_ERR_CUDA;
        grid.y = 1;              // one CTA per frame: warp 0 drives the TMA engine, all eight warps build the reference rows
    }
    rows_gather_kernel<TF, TB><<<grid, 256, sm, st>>>(*a, rp, bulk);
    return TSCD_OK;
}

template <typename TF>
static int launch_gather_tb(const tscd_gather_args* a, dim3 grid, cudaStream_t st) {
    const int rp = fused_rows_pitch(a->anchors, a->reg, a->obj, a->cls, a->num_classes, a->head_dtype, false);
    const int bulk = gather_bulk_ok(a) ? 1 : 0;
    switch (a->bank_dtype) {
        case TSCD_F32: return launch_gather_kernel<TF, float>(a, grid, rp, 0, st);
        case TSCD_F16: return launch_gather_kernel<TF, __half>(a, grid, rp, bulk, st);
        case TSCD_BF16: return launch_gather_kernel<TF, __nv_bfloat16>(a, grid, rp, bulk, st);
        default: return TSCD_ERR_UNSUPPORTED;
    }
    return TSCD_OK;
}

// Prefix offsets of the LOCAL rows (first L frames of every clip) from the per-frame counts: lrow_off[b*L + f].
// One CTA; replaces a handful of tiny framework kernels (slice, cumsum, copy) on the critical path.
__global__ void local_offsets_kernel(int B, int F, int L, const int32_t* sel_count, int32_t* lrow_off) {
    __shared__ int scratch[40];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int n = B * L;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int c = 0;
        if (i < n) { const int b = i / L, f = i - b * L; c = sel_count[b * F + f]; }
        int tot;
        const int ex = block_excl_scan(c, scratch, &tot);
        const int carry = carry_s;
        if (i < n) lrow_off[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) lrow_off[n] = carry_s;
}

}  // namespace tscd

extern "C" int tscd_gather(const tscd_gather_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->num_frames < 0 || a->num_classes <= 0 || a->max_keep <= 0 || a->feat_dim <= 0) return TSCD_ERR_INVALID_ARG;
    if (a->head_dtype != TSCD_F32 && a->head_dtype != TSCD_F16) return TSCD_ERR_UNSUPPORTED;
    if (a->num_frames == 0) return TSCD_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    count_scan_kernel<<<1, 1024, 0, st>>>(a->num_frames, a->use_keep, a->max_keep, a->cand_count, a->keep_count,
                                          a->sel_count, a->row_off);
    TSCD_CUDA_CHECK_LAUNCH();
    // one warp per kept row where possible (a row costs 3-4 dependent global round trips; rows are independent)
    int ysplit = (a->max_keep + 7) / 8;
    if (ysplit > 16) ysplit = 16;
    dim3 grid(a->num_frames, ysplit);
    int rc;
    switch (a->feat_dtype) {
        case TSCD_F32: rc = launch_gather_tb<float>(a, grid, st); break;
        case TSCD_F16: rc = launch_gather_tb<__half>(a, grid, st); break;
        case TSCD_BF16: rc = launch_gather_tb<__nv_bfloat16>(a, grid, st); break;
        default: return TSCD_ERR_UNSUPPORTED;
    }
    if (rc != TSCD_OK) return rc;
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}

extern "C" int tscd_local_offsets(const tscd_local_offsets_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->B <= 0 || a->F <= 0 || a->L <= 0 || a->L > a->F || !a->sel_count || !a->lrow_off) return TSCD_ERR_INVALID_ARG;
    local_offsets_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a->B, a->F, a->L, a->sel_count, a->lrow_off);
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}
