// Long-clip mode (BASELINE.json configs[3]): one clip sharded by frame over the ranks.  Every rank
// selects / gathers ITS frames; the attention of a rank's local frames needs the bank rows of ALL global frames as keys and
// values, so the ranks exchange those rows with ONE all-gather (NCCL over NVLink/NVSwitch, issued by the host through
// torch.distributed) of one packed buffer per rank:
//
//   [ header: 32 ints -- counts of the rank's G_r global frames ][ G_r * kmax rows of (cls 256 | reg 256 | score, pad) ]
//
// bank_pack builds that buffer from the rank's packed bank (rows of a frame padded to kmax so the layout is static: the host
// never needs the data-dependent counts), bank_unpack builds the virtual clip  [own local frames | every rank's global frames,
// rank-major]  from the gathered buffers: per-frame counts, prefix offsets and the compacted 16-bit rows, ready for
// tscd_qkv_project / the attention kernels.  Edge features, objectness and boxes of the global rows are not exchanged: only the
// local rows use them (CAFM, final expansion).
#include "common.cuh"

namespace tscd {

constexpr int kXhdrInts = 256;                      // header ints (counts of up to 256 global frames per rank)
constexpr int kXrowBytes = 2 * 256 * 2 + 16;        // cls | reg (16-bit) | score fp32 + pad

__global__ void __launch_bounds__(256) bank_pack_kernel(const tscd_bank_pack_args a) {
    const int gf = blockIdx.x;                      // global frame of this rank
    const int f = a.n_local_frames + gf;
    const int n = min(a.sel_count[f], a.kmax);
    const int r0 = a.row_off[f];
    unsigned char* out = reinterpret_cast<unsigned char*>(a.send);
    if (gf == 0)
        for (int i = threadIdx.x; i < kXhdrInts; i += blockDim.x)
            reinterpret_cast<int32_t*>(out)[i] = i < a.n_global_frames ? min(a.sel_count[a.n_local_frames + i], a.kmax) : 0;
    unsigned char* rows = out + kXhdrInts * 4 + (size_t)gf * a.kmax * kXrowBytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < n; j += nw) {
        const uint4* c = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.bank_cls) + (size_t)(r0 + j) * 512);
        const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.bank_reg) + (size_t)(r0 + j) * 512);
        uint4* d = reinterpret_cast<uint4*>(rows + (size_t)j * kXrowBytes);
        d[lane] = c[lane];
        d[32 + lane] = r[lane];
        if (lane == 0) d[64] = make_uint4(__float_as_uint(a.bank_score[r0 + j]), 0u, 0u, 0u);
    }
}

// One CTA: counts + prefix offsets of the virtual clip.
__global__ void __launch_bounds__(1024) bank_unpack_offsets_kernel(const tscd_bank_unpack_args a) {
    __shared__ int scan[40];
    const int Fv = a.n_local_frames + a.world * a.n_global_frames;
    int carry = 0;
    for (int f0 = 0; f0 < Fv; f0 += blockDim.x) {
        const int f = f0 + threadIdx.x;
        int c = 0;
        if (f < a.n_local_frames) c = min(a.sel_count[f], a.kmax);
        else if (f < Fv) {
            const int r = (f - a.n_local_frames) / a.n_global_frames, g = (f - a.n_local_frames) % a.n_global_frames;
            c = reinterpret_cast<const int32_t*>(reinterpret_cast<const unsigned char*>(a.recv) + (size_t)r * a.rank_bytes)[g];
        }
        int tot;
        const int ex = block_excl_scan(c, scan, &tot);
        if (f < Fv) { a.v_count[f] = c; a.v_row_off[f] = carry + ex; }
        carry += tot;
    }
    if (threadIdx.x == 0) a.v_row_off[Fv] = carry;
}

__global__ void __launch_bounds__(256) bank_unpack_rows_kernel(const tscd_bank_unpack_args a) {
    const int f = blockIdx.x;
    const int n = a.v_count[f];
    const int d0 = a.v_row_off[f];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (f < a.n_local_frames) {                    // own local frame: rows already packed in the rank's bank
        const int s0 = a.row_off[f];
        for (int j = warp; j < n; j += nw) {
            reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(a.v_bank_cls) + (size_t)(d0 + j) * 512)[lane] =
                reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.bank_cls) + (size_t)(s0 + j) * 512)[lane];
            reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(a.v_bank_reg) + (size_t)(d0 + j) * 512)[lane] =
                reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.bank_reg) + (size_t)(s0 + j) * 512)[lane];
            if (a.v_bank_edge)
                reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(a.v_bank_edge) + (size_t)(d0 + j) * 512)[lane] =
                    reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.bank_edge) + (size_t)(s0 + j) * 512)[lane];
            if (lane == 0) a.v_bank_score[d0 + j] = a.bank_score[s0 + j];
        }
        return;
    }
    const int r = (f - a.n_local_frames) / a.n_global_frames, g = (f - a.n_local_frames) % a.n_global_frames;
    const unsigned char* rows = reinterpret_cast<const unsigned char*>(a.recv) + (size_t)r * a.rank_bytes + kXhdrInts * 4 +
                                (size_t)g * a.kmax * kXrowBytes;
    for (int j = warp; j < n; j += nw) {
        const uint4* s = reinterpret_cast<const uint4*>(rows + (size_t)j * kXrowBytes);
        reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(a.v_bank_cls) + (size_t)(d0 + j) * 512)[lane] = s[lane];
        reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(a.v_bank_reg) + (size_t)(d0 + j) * 512)[lane] = s[32 + lane];
        if (lane == 0) a.v_bank_score[d0 + j] = __uint_as_float(s[64].x);
    }
}

}  // namespace tscd

extern "C" int64_t tscd_bank_pack_bytes(int32_t n_global_frames, int32_t kmax) {
    if (n_global_frames < 0 || n_global_frames > tscd::kXhdrInts || kmax <= 0) return -1;
    return (int64_t)tscd::kXhdrInts * 4 + (int64_t)n_global_frames * kmax * tscd::kXrowBytes;
}

extern "C" int tscd_bank_pack(const tscd_bank_pack_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->n_global_frames <= 0 || a->n_global_frames > kXhdrInts || a->kmax <= 0 || a->n_local_frames < 0) return TSCD_ERR_INVALID_ARG;
    if (a->dtype != TSCD_F16 && a->dtype != TSCD_BF16) return TSCD_ERR_UNSUPPORTED;
    bank_pack_kernel<<<a->n_global_frames, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}

extern "C" int tscd_bank_unpack(const tscd_bank_unpack_args* a, void* stream) {
    using namespace tscd;
    if (!a || a->world <= 0 || a->n_global_frames <= 0 || a->n_global_frames > kXhdrInts || a->kmax <= 0 || a->n_local_frames < 0)
        return TSCD_ERR_INVALID_ARG;
    if (a->rank_bytes < tscd_bank_pack_bytes(a->n_global_frames, a->kmax)) return TSCD_ERR_INVALID_ARG;
    if (a->dtype != TSCD_F16 && a->dtype != TSCD_BF16) return TSCD_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    bank_unpack_offsets_kernel<<<1, 1024, 0, st>>>(*a);
    TSCD_CUDA_CHECK_LAUNCH();
    bank_unpack_rows_kernel<<<a->n_local_frames + a->world * a->n_global_frames, 256, 0, st>>>(*a);
    TSCD_CUDA_CHECK_LAUNCH();
    return TSCD_OK;
}
