// Tensor-core convolution for sm_100a: tcgen05.mma (kind::f16, FP16 operands, FP32 accumulation in
// TMEM), operands staged in shared memory by the TMA bulk-copy engine (cp.async.bulk + mbarrier),
// warp-specialised persistent CTAs, fused epilogue (bias, activation, BatchNorm affine, residual,
// per-(sample,channel) scale) that writes the next layer's operand layout directly.
//
// ---------------------------------------------------------------------------------------------
// Formulation ("row-group transposed implicit GEMM")
//   The layers have only <= 32 output channels, so the usual pixels-as-M / Cout-as-N GEMM would be
//   bound by shared-memory operand bandwidth (N = 32 -> A re-read every 16 cycles).  Instead the
//   WEIGHTS are the M operand and the PIXELS the N operand:
//       D[(r, co), x] = sum_{rho, dx, ci} A_{rho,dx}[(r, co), ci] * X[y0 + rho - pad, x + dx - pad, ci]
//   with M = 128 = 4 output rows r x 32 channels co, N = up to 256 pixels of one image row segment,
//   K = 16 input channels per MMA.  For input row rho (0 <= rho < kh+3) and column tap dx the A tile
//   is  A[(r,co)] = W[dy = rho - r, dx, ci, co]  (zero when dy is outside [0,kh)).  With the M rows
//   ordered m = (3-r)*32 + co and the weights packed per (ci-chunk, dx) as Wd[z = dy+3][co] with 3
//   zero z-rows on both ends, that A tile is the CONTIGUOUS window of Wd starting at row rho*32, so
//   one packed array per (chunk, dx) serves all kh+3 input rows by moving the descriptor start
//   address.  Cost: (kh+3)/kh more MMAs than the minimum, in exchange for a 128x256x16 MMA shape
//   (96 B/cycle of shared-memory operand traffic, inside the 128 B/cycle budget), no im2col, no
//   epilogue reduction, and an accumulator whose lanes are (row, channel) and columns are pixels.
//
// Data layout ("BLK8"): activations are fp16 [B][C/8][H+14][W+14][8]: channel planes of 8, a 7-pixel
//   halo materialised in global memory (zero for CONSTANT padding, mirrored by pcnn_blk8_halo_fill
//   for SYMMETRIC), 16 bytes per pixel per plane.  A row segment of one plane is contiguous, so every
//   operand load is a plain 1-D bulk copy and the kernel is padding-agnostic.  In shared memory the
//   same bytes are a K-major no-swizzle UMMA operand: rows (pixels) 16 B apart, 8-row core matrices
//   128 B apart (SBO), the two 8-channel K halves one plane apart (LBO); a column tap is a +16 B
//   start-address shift.
//
// Warp roles (352 threads): 0 = input-row producer, 1 = weight producer, 2 = MMA issuer (+TMEM
//   alloc), 3..10 = epilogue (TMEM lane quarter = warp % 4, two warps per quarter split the columns).
//   Two 256-column accumulators in TMEM let the epilogue of tile t overlap the MMAs of tile t+1.
#include <cuda_fp16.h>
#include <climits>
#include <cstdlib>
#include <cuda_fp8.h>

#include "pcnn_common.cuh"

namespace pcnn {
namespace tc {

constexpr int HALO = 7;            // materialised halo of the BLK8 layout (kernel sizes up to 15)
// The M = 128 accumulator rows are RT output rows x CP channel slots, CP = 32, 16 or 8 (smallest >= Cout):
// narrow layers trade channel padding for more output rows per tile, i.e. fewer MMAs per output row
// ((kh+RT-1)*kw per RT rows) and an epilogue whose lanes all carry live channels.
constexpr int M_TILE = 128;
// output rows per tile for CP channel slots: 128/CP, and 5 (120 of the 128 accumulator rows) for CP = 24
__host__ __device__ constexpr int rows_per_tile(int cp) { return cp == 24 ? 5 : M_TILE / cp; }
// z-rows of the packed weights per (chunk, tap, K-half): kh taps + RT-1 zero rows on each side; CP = 24 gets one more
// so that the 128-row window of the last input row (8 rows past 5 x 24) stays inside the stage
__host__ __device__ constexpr int packed_zrows(int kh, int cp) { return kh + 2 * (rows_per_tile(cp) - 1) + (cp == 24 ? 1 : 0); }
constexpr int MAX_EPI_WARPS = 8;               // 12 for k <= 5 (second kernel build): those epilogues are latency-bound
constexpr int CHUNK_PX = 16;                   // accumulator columns per epilogue pass
constexpr int STAGE_WARP = CHUNK_PX * 128;     // per-warp transpose buffer: 16 pixels x 32 words
constexpr float LO_SCALE = 2048.f;        // 2^11: brings the fp16 rounding remainder into e4m3's range (mode 3)
constexpr unsigned long long SPIN_LIMIT_NS = 4000000000ull;   // a stuck pipeline traps instead of hanging the GPU

struct Params {
    const __half* in;        // BLK8 [B][c8_in][Hp][P][8]  (hi part)
    const __half* in_lo;     // lo part (split precision) or null
    const __half* wpack;     // [nsplit][C16][kw][2][(kh+2(RT-1))*CP][8]  (hi image, then lo image)
    const float* bias;       // [Cout] or null
    const float* bn_scale;   // [Cout] or null
    const float* bn_shift;
    const __half* residual;  // BLK8 like out, or null
    const __half* residual_lo;
    const float* out_scale;  // [B][Cout] or null
    __half* out;             // BLK8 [B][c8_out][Hp][P][8]
    __half* out_lo;          // lo part of the output (split precision) or null
    float acc_scale;         // exact power of two undoing the weight pre-scaling (applied to the accumulator)
    int nsplit;              // precision mode: 1 single FP16 pass; 2 hi/lo fp16 split (3 MMAs per term);
                             // 3 fp16 main pass + ONE e4m3 K=32 MMA for both correction terms (in_lo/out_lo/residual_lo
                             // then point at the fp8 "q" buffers: planes 2c = e4m3(x), 2c+1 = e4m3((x - hi) * 2^11))
    int kpass;               // MMA passes per 16-channel chunk: 1 (fp16 only), 2 (fp16 + e4m3 correction), 3 (hi/lo split).
                             // kpass = 1 with nsplit = 3 tensors: a layer whose correction pass is skipped
    int nv;                  // virtual K-chunks = c16 * kpass
    int B, H, W, Hp, P;
    int c8_in, c8_out, c8_res;
    int c16;                 // input-channel chunks of 16
    int cout;                // true output channels
    int kh, kw, pad;
    int act;
    int n_tile;              // MMA N (multiple of 16, <= 256)
    int tiles_x, tiles_y, num_tiles;
    int row_slots;           // ring of input-row slots (>= kh+RT-1)
    int w_stages;
    int halo_sym;            // 1: the epilogue also writes the SYMMETRIC (edge-repeating mirror) halo ring of width 7 of `out`
    int n_epi;               // active epilogue warps: 8, or 4 when shared memory is needed for operands (large k)
    int debug;               // PCNN_TC_DEBUG bit 0: epilogue skips math and stores (timing experiments only)
    int w_resident;          // 1: all weight stages of a tile fit in shared memory -> loaded once per CTA, reused by every tile
    int pair_tail;           // 1: the last 16-channel chunk holds ONE live 8-channel plane (ceil(Cin/8) odd): its fp16 passes put two
                             // adjacent column taps into the two K halves (B: K-half stride = one pixel) -> (kw+1)/2 MMAs per row, not kw
    int npairs;              // (kw+1)/2 weight stages of a paired chunk
    int res_tail_pl;         // mode 3: plane of the residual tensor holding a lone live octet (ceil(Cres/8) odd), else -1
    int rw;                  // 1: "row weights" (pcnn_conv2d_tc_rowweights): the input is ONE row (a 1-D signal, H = 1) and every output
                             // row has its own kw x Cin x Cout weights: wpack = [c16][kw][2][T slots][CP][8], slot t = tile*RT + (RT-1-r);
                             // per tile one input row and c16*kw MMAs (no row taps, no zero z-rows)
    int rw_tslots;           // T
    uint32_t rowplane_bytes; // bytes of one plane of one row window in smem (multiple of 128)
    uint32_t row_copy_bytes; // (n_tile + kw - 1) * 16
    uint32_t wstage_bytes;   // 2 * (kh+2(RT-1)) * CP * 16
    uint32_t idesc;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_try_wait(bar, parity)) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > SPIN_LIMIT_NS) {
            printf("pcnn conv_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (cute::elect_one_sync)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// tanh to ~1e-7 absolute without the slow libm path: odd polynomial near 0, 1 - 2/(e^{2x}+1) elsewhere
__device__ __forceinline__ float tanh_fast(float x) {   // branch-free: both forms are a handful of instructions
    const float ax = fabsf(x), x2 = x * x;
    const float small = x * fmaf(x2, fmaf(x2, 0.13333334f, -0.33333334f), 1.0f);
    const float e = __expf(2.0f * fminf(ax, 15.0f));
    const float big = copysignf(1.0f - __fdividef(2.0f, e + 1.0f), x);
    return ax < 0.08f ? small : big;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same shape with e4m3 operands: K = 32 per instruction at the same cycle cost (2x the fp16 rate)
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint8_t to_e4m3(float f) { return (uint8_t)__nv_cvt_float_to_fp8(f, __NV_SATFINITE, __NV_E4M3); }
__device__ __forceinline__ float from_e4m3(uint8_t b) { return __half2float(__half(__nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3))); }
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// 16 consecutive accumulator columns of this thread's TMEM lane; completes at tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float2 e4m3x2_to_float2(uint32_t two_bytes) {
    const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)two_bytes, __NV_E4M3);
    return __half22float2(__half2(h));
}
__device__ __forceinline__ uint32_t float2_to_e4m3x2(float lo, float hi) {   // byte 0 = e4m3(lo), byte 1 = e4m3(hi)
    return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ float2 bits_to_float2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }

// ---------------------------------------------------------------- the kernel
// CP = channel slots per output row of the M operand (32, 24, 16 or 8); RT = rows_per_tile(CP) output rows per tile.
// MODE = precision mode of the tensors (1, 2, 3): compile-time in the epilogue, which is instruction-bound on narrow layers.
// NEPI = epilogue warps the build allows (8 or 12: the register budget per thread follows from it).
template <int CP, int MODE, int NEPI>
__global__ void __launch_bounds__((3 + NEPI) * 32, 1) conv_tc_kernel(const Params p) {
    constexpr int RT = rows_per_tile(CP);      // output rows per tile
    constexpr int ZPAD = RT - 1;         // zero z-rows on each side of the packed weights
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve-up: [row slots][weight stages][epilogue transpose buffers n_epi x 2 KB][barriers][tmem ptr]
    uint8_t* s_rows = smem;
    const uint32_t row_slot_bytes = 2 * p.rowplane_bytes;
    uint8_t* s_w = s_rows + (size_t)p.row_slots * row_slot_bytes;
    uint8_t* s_stage = s_w + (size_t)p.w_stages * p.wstage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + p.n_epi * STAGE_WARP);
    uint64_t* row_full = bars;
    uint64_t* row_empty = row_full + p.row_slots;
    uint64_t* w_full = row_empty + p.row_slots;
    uint64_t* w_empty = w_full + p.w_stages;
    uint64_t* acc_full = w_empty + p.w_stages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.rw ? 1 : p.kh + ZPAD;   // input rows per tile
    // rows travel in two barrier groups per chunk: R/2 each when R is even (row_slots is then a multiple of R/2 and a
    // group never wraps), (R+1)/2 and R/2 when it is odd (CP = 24; row_slots is then a multiple of R)
    const int GR0 = (R + 1) / 2, GR1 = R / 2;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.row_slots; ++i) { mbar_init(row_full + i, 1); mbar_init(row_empty + i, 1); }
        for (int i = 0; i < p.w_stages; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, p.n_epi); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM: 512 columns = two 256-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // the weight stages start as zeros: the producer only ever copies the live z-rows into them
    for (uint32_t i = threadIdx.x; i < (uint32_t)p.w_stages * (p.wstage_bytes >> 4); i += blockDim.x)
        reinterpret_cast<uint4*>(s_w)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core's reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= input-row producer =================
        // converged warp, one elected lane issues the bulk copies; ring state advances by compare (no div/mod)
        const bool leader = elect_one();
        const size_t plane_elems = (size_t)(p.rw ? 1 + 2 * HALO : p.Hp) * p.P * 8;    // row-weights mode: the input tensor has H = 1
        uint32_t slot = 0, ph = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            const int tx = t % p.tiles_x;
            const int ty = (t / p.tiles_x) % p.tiles_y;
            const int b = t / (p.tiles_x * p.tiles_y);
            const int x0 = tx * p.n_tile, y0 = ty * RT;
            const int col0 = x0 + HALO - p.pad, prow0 = y0 + HALO - p.pad;
            for (int v = 0; v < p.nv; ++v) {
                // split precision: per chunk c the passes are (x_hi,W_hi), (x_hi,W_lo), (x_lo,W_hi); mode 3: (x_hi,W_hi), (q,Wq)
                const int c = (p.kpass == 3) ? v / 3 : (p.kpass == 2 ? v >> 1 : v);
                const __half* inp = ((p.kpass == 3 && (v % 3) == 2) || (p.kpass == 2 && (v & 1))) ? p.in_lo : p.in;
                const __half* base = inp + ((size_t)b * p.c8_in + 2 * c) * plane_elems + (size_t)col0 * 8;
                const bool paired = p.pair_tail && c == p.c16 - 1;   // only plane 2c travels (fp16 and e4m3 passes alike)
                // rows travel in two groups per chunk (R is even); one mbarrier pair per group, at the group's first slot
                for (int grp = 0, rho = 0; grp < 2; ++grp) {
                    const int GR = grp ? GR1 : GR0;
                    if (GR == 0) continue;                 // row-weights mode: a single input row, one group
                    mbar_wait(row_empty + slot, ph ^ 1);
                    if (leader) {
                        mbar_expect_tx(row_full + slot, (uint32_t)GR * (paired ? 1u : 2u) * p.row_copy_bytes);
                        for (int i = 0; i < GR; ++i, ++rho) {
                            const __half* src = base + (size_t)(p.rw ? HALO : min(prow0 + rho, p.Hp - 1)) * p.P * 8;
                            const uint32_t dst = smem_u32(s_rows) + (slot + i) * row_slot_bytes;
                            bulk_copy_g2s(dst, src, p.row_copy_bytes, row_full + slot);
                            if (!paired) bulk_copy_g2s(dst + p.rowplane_bytes, src + plane_elems, p.row_copy_bytes, row_full + slot);
                        }
                    } else {
                        rho += GR;
                    }
                    slot += GR;
                    if (slot == (uint32_t)p.row_slots) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= weight producer =================
        const bool leader = elect_one();
        const uint32_t w_plane = p.wstage_bytes >> 1, w_zoff = (uint32_t)ZPAD * CP * 16u, w_live = (uint32_t)p.kh * CP * 16u;
        uint32_t st = 0, ph = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            if (p.w_resident && t != (int)blockIdx.x) break;    // resident weights: one pass fills every stage
            if (p.rw) {
                // row weights: per (chunk, column tap) the 128 M rows of THIS tile: slots [ty*RT, ...) of each K half
                const int ty = (t / p.tiles_x) % p.tiles_y;
                const uint32_t khalf_g = (uint32_t)p.rw_tslots * CP * 16u;       // K-half stride in global memory
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)ty * RT * CP * 16u;
                for (int s2 = 0; s2 < p.c16 * p.kw; ++s2, src += 2u * khalf_g) {
                    mbar_wait(w_empty + st, ph ^ 1);
                    if (leader) {
                        mbar_expect_tx(w_full + st, 4096u);
                        const uint32_t dst = smem_u32(s_w) + st * p.wstage_bytes;
                        bulk_copy_g2s(dst, src, 2048u, w_full + st);
                        bulk_copy_g2s(dst + 2048u, src + khalf_g, 2048u, w_full + st);
                    }
                    if (++st == (uint32_t)p.w_stages) { st = 0; ph ^= 1; }
                }
                continue;
            }
            for (int v = 0; v < p.nv; ++v) {
                const int c = (p.kpass == 3) ? v / 3 : (p.kpass == 2 ? v >> 1 : v);
                const int wsel = ((p.kpass == 3 && (v % 3) == 1) || (p.kpass == 2 && (v & 1))) ? 1 : 0;   // second weight image
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wpack) + ((size_t)wsel * p.c16 + c) * p.kw * p.wstage_bytes;
                // a paired chunk has (kw+1)/2 stages: K-half 0 / 1 = column taps (2s, 2s+1)
                const int nst = (p.pair_tail && c == p.c16 - 1) ? p.npairs : p.kw;
                for (int dx = 0; dx < nst; ++dx, src += p.wstage_bytes) {
                    mbar_wait(w_empty + st, ph ^ 1);
                    if (leader) {
                        // only the kh live z-rows of each K-half travel: the zero rows around them were written once at
                        // kernel start and no copy ever touches them (44..71 % of a stage's bytes)
                        mbar_expect_tx(w_full + st, 2u * w_live);
                        const uint32_t dst = smem_u32(s_w) + st * p.wstage_bytes + w_zoff;
                        bulk_copy_g2s(dst, src + w_zoff, w_live, w_full + st);
                        bulk_copy_g2s(dst + w_plane, src + w_plane + w_zoff, w_live, w_full + st);
                    }
                    if (++st == (uint32_t)p.w_stages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ================= MMA issuer =================
        // The whole warp stays converged (all loop state is warp-uniform and lives in uniform
        // registers); one elected lane issues tcgen05.mma / tcgen05.commit.  Everything per MMA is
        // incremental: descriptor low words advance by constants, ring slots wrap by compare.
        {
            const uint32_t a_lbo = p.rw ? 2048u : (uint32_t)packed_zrows(p.kh, CP) * (CP * 16u);   // K-half (plane) stride of the packed weights
            const uint32_t a_hi = (uint32_t)(make_desc(0, a_lbo, 128u) >> 32);
            const uint32_t b_hi = (uint32_t)(make_desc(0, p.rowplane_bytes, 128u) >> 32);
            const uint32_t a_lo_lbo = ((a_lbo >> 4) & 0x3FFF) << 16, b_lo_lbo = ((p.rowplane_bytes >> 4) & 0x3FFF) << 16;
            const uint32_t rows_base16 = smem_u32(s_rows) >> 4, w_base16 = smem_u32(s_w) >> 4;
            const uint32_t slot16 = row_slot_bytes >> 4, wstage16 = p.wstage_bytes >> 4;
            const uint32_t nslots = p.row_slots, nwst = p.w_stages;
            const bool leader = elect_one();
            uint32_t slot0 = 0, slot0_ph = 0;    // ring position/phase of the current chunk's row 0
            uint32_t wst = 0, wph = 0, it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
                const uint32_t acc = it & 1, acc_ph = (it >> 1) & 1;
                mbar_wait(acc_empty + acc, acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                uint32_t accum = 0;
                for (int c = 0; c < p.nv; ++c) {
                    const bool f8 = (p.kpass == 2) && (c & 1);    // correction pass: e4m3 operands, K = 32
                    // paired chunk (one live 8-channel plane): the K halves are column taps (dx, dx+1) of that plane, i.e. the
                    // B descriptor's K-half stride is one pixel (16 B); the last stage is taps (kw-2, kw-1) with a zero first half
                    // (e4m3 pass: the lone q plane carries [e4m3(x) | e4m3(x_lo)] of its 8 channels per pixel, so the same pairing holds)
                    const bool paired = p.pair_tail && ((p.kpass == 3) ? c / 3 : (p.kpass == 2 ? c >> 1 : c)) == p.c16 - 1;
                    const int nst = paired ? p.npairs : p.kw;
                    const uint32_t b_lbo = paired ? (1u << 16) : b_lo_lbo;
                    // the chunk's rows sit in two groups of GR slots: [slot0, +GR) and the next group (which may wrap to 0)
                    const uint32_t slotA = slot0, phA = slot0_ph;
                    uint32_t slotB = slot0 + GR0, phB = slot0_ph;
                    if (slotB == nslots) { slotB = 0; phB ^= 1; }
                    for (int st = 0; st < nst; ++st) {
                        const int dx = paired ? min(2 * st, p.kw - 2) : st;   // first column tap of the stage
                        mbar_wait(w_full + wst, p.w_resident ? 0u : wph);   // resident stages complete once and stay
                        tc_fence_after();
                        uint32_t a_lo = ((w_base16 + wst * wstage16) & 0x3FFF) | a_lo_lbo;
                        const bool first_dx = (st == 0), last_dx = (st == nst - 1);
#pragma unroll 1
                        for (int grp = 0; grp < 2; ++grp) {
                            const uint32_t slot = grp ? slotB : slotA;
                            const int GR = grp ? GR1 : GR0;
                            if (GR == 0) continue;
                            if (first_dx) { mbar_wait(row_full + slot, grp ? phB : phA); tc_fence_after(); }
                            if (leader) {
                                uint32_t b_lo = ((rows_base16 + slot * slot16 + dx) & 0x3FFF) | b_lbo;
                                if (!f8) {
#pragma unroll 3
                                    for (int i = 0; i < GR; ++i) {
                                        tc_mma_f16(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, p.idesc, accum | (uint32_t)(i | grp | st));
                                        a_lo += CP; b_lo += slot16;
                                    }
                                } else {
#pragma unroll 3
                                    for (int i = 0; i < GR; ++i) {
                                        tc_mma_f8(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, p.idesc, 1u);
                                        a_lo += CP; b_lo += slot16;
                                    }
                                }
                                if (last_dx) tc_commit(row_empty + slot);   // the whole group is free again
                            } else {
                                a_lo += (uint32_t)CP * GR;
                            }
                        }
                        accum = 1;
                        if (leader && !p.w_resident) tc_commit(w_empty + wst);
                        if (++wst == nwst) { wst = 0; wph ^= 1; }
                    }
                    slot0 += R;
                    if (slot0 >= nslots) { slot0 -= nslots; slot0_ph ^= 1; }
                }
                if (leader) tc_commit(acc_full + acc);
                __syncwarp();
            }
        }
    } else if (warp - 3 < p.n_epi) {
        // ================= epilogue (warps 3..3+n_epi) =================
        // Two thread roles per warp, joined by a 16 px x 32 word transpose buffer (XOR-swizzled 16-byte chunks,
        // conflict-free both ways):
        //   "lane side"  : thread = TMEM lane m = 32q + lane = (row slot, channel), 16 consecutive pixels
        //                  (what tcgen05.ld delivers); does bias / activation / BN / residual add / rounding.
        //   "pixel side" : thread = (pixel px, octet pair oc): 2 x 8 consecutive M rows of one pixel = the
        //                  16-byte units of the BLK8 layout; does all global loads (residual) and stores,
        //                  each warp instruction covering 16 consecutive pixels x 16 B = 256 contiguous bytes.
        // One 32-bit word per (pixel, M row) carries every output form: [fp16 hi | e4m3(x) | e4m3(lo*2^11)]
        // (mode 3), [hi | lo] (mode 2) or [hi] (mode 1).
        const int ew = warp - 3;
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int part = ew >> 2, nparts = p.n_epi >> 2;
        const int m = 32 * q + lane;
        const int co = m % CP;                      // CP = 24: rows 120..127 are never stored
        uint32_t* stage = reinterpret_cast<uint32_t*>(s_stage + ew * STAGE_WARP);
        const bool live = co < p.cout && m < RT * CP;
        const float bias = (p.bias && live) ? p.bias[co] : 0.f;
        const float bns = (p.bn_scale && live) ? p.bn_scale[co] : 1.f;
        const float bnt = (p.bn_shift && live) ? p.bn_shift[co] : 0.f;
        const int planes_out = (p.cout + 7) / 8;
        // mode 3 "tail" plane: with an odd number of live 8-channel planes the last one is alone in its 16-channel group; its
        // q plane then holds [e4m3(x) x 8 | e4m3(x_lo * 2^11) x 8] per pixel (one plane instead of two half-empty ones), which
        // is what lets the consumer pair column taps in the e4m3 pass as well
        const int tail_pl = (planes_out & 1) ? planes_out - 1 : -1;
        const int nchunks = p.n_tile / CHUNK_PX;
        const int ch_begin = (part * nchunks) / nparts, ch_end = ((part + 1) * nchunks) / nparts;   // balanced column split
        const int act = p.act;
        constexpr int mode = MODE;
        const bool affine = (p.bn_scale != nullptr) || (p.out_scale != nullptr);
        const float asc = p.acc_scale;
        const bool has_res = p.residual != nullptr;
        // pixel-side identity
        const int px = lane & 15, oc = lane >> 4;
        int o_r[2], o_pl[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int m0 = 8 * (4 * q + 2 * oc + e);
            o_r[e] = (RT - 1) - m0 / CP;
            o_pl[e] = (m0 % CP) >> 3;
            if (m0 >= RT * CP) o_pl[e] = 1 << 20;    // CP = 24: the last octet (rows 120..127) belongs to no output row
        }
        // swizzled word offsets: word (pixel j, M row lane) and the pixel side's four 16-byte chunks
        const uint32_t lane_word = (uint32_t)(lane & 3);
        const uint32_t lane_chunk = (uint32_t)(lane >> 2);
        uint32_t* const px_row = stage + px * 32;
        const uint32_t px_sw = (uint32_t)(px & 7);
        const size_t plane_px = (size_t)p.Hp * p.P;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int tx = t % p.tiles_x;
            const int ty = (t / p.tiles_x) % p.tiles_y;
            const int b = t / (p.tiles_x * p.tiles_y);
            const int x0 = tx * p.n_tile;
            const uint32_t acc = it & 1, acc_ph = (it >> 1) & 1;
            // fold the per-(sample,channel) scale into the BN affine: (a*s + t) * o = a*(s*o) + t*o
            const float osc = (p.out_scale && live) ? p.out_scale[(size_t)b * p.cout + co] : 1.f;
            const float mul = bns * osc, add = bnt * osc;
            // pixel side: per octet, pixel index of (row y, x = x0 + px) within a plane, and validity
            size_t o_pix[2], o_off[2], r_off[2];
            bool o_ok[2];
            int o_dy[2][2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int y = ty * RT + o_r[e];
                o_ok[e] = (y < p.H) && (o_pl[e] < planes_out);
                o_pix[e] = (size_t)(y + HALO) * p.P + (x0 + px + HALO);
                // byte offsets of this thread's 16-byte unit in `out` (and, identically, in the lo / q buffer: every
                // plane is 16 B per pixel) and in the residual; per chunk only the column advances
                o_off[e] = (((size_t)b * p.c8_out + o_pl[e]) * plane_px + o_pix[e]) * 16;
                r_off[e] = (((size_t)b * p.c8_res + o_pl[e]) * plane_px + o_pix[e]) * 16;
                o_dy[e][0] = (p.halo_sym && y < HALO) ? -(2 * y + 1) : 0;                       // row displacement of the top mirror
                o_dy[e][1] = (p.halo_sym && y >= p.H - HALO && y < p.H) ? 2 * (p.H - y) - 1 : 0;  // ... and of the bottom mirror
            }
            uint4 rh[2], rl[2];
            auto load_residual = [&](int ch) {
                const int xo = ch * CHUNK_PX;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    rh[e] = make_uint4(0, 0, 0, 0); rl[e] = make_uint4(0, 0, 0, 0);
                    if (o_ok[e] && x0 + xo + px < p.W) {
                        const size_t off = r_off[e] + (size_t)xo * 16;
                        rh[e] = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.residual) + off);
                        if (mode == 2) rl[e] = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.residual_lo) + off);
                        if (mode == 3) {
                            // remainder plane 2g+1 of this octet's 16-channel group: one plane up for an even plane, same plane otherwise
                            // (tail plane: bytes 8..15 of the same plane)
                            const uint2 t2 = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(p.residual_lo) + off +
                                (((o_pl[e] & 1) || o_pl[e] == p.res_tail_pl) ? (size_t)8 : plane_px * 16));
                            rl[e].x = t2.x; rl[e].y = t2.y;
                        }
                    }
                }
            };
            if (has_res && ch_begin < ch_end) load_residual(ch_begin);   // does not depend on the accumulator
            mbar_wait(acc_full + acc, acc_ph);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
            // the accumulator columns of chunk ch+1 are requested as soon as chunk ch has arrived, so the TMEM read latency
            // hides behind the math and stores of chunk ch (it was the critical path of the narrow, small-k layers)
            uint32_t vn[16];
            if (ch_begin < ch_end) tmem_ld16_issue(taddr0 + ch_begin * CHUNK_PX, vn);
            for (int ch = ch_begin; ch < ch_end; ++ch) {
                uint32_t v[16];
                float r[16];
                if (has_res) {
                    // pixel side: residual -> fp32 -> transpose buffer
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const uint32_t hw[4] = {rh[e].x, rh[e].y, rh[e].z, rh[e].w};
                        const uint32_t lw[4] = {rl[e].x, rl[e].y, rl[e].z, rl[e].w};
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 a = bits_to_float2(hw[i]);
                            f[2 * i] = a.x; f[2 * i + 1] = a.y;
                            if (mode == 2) {
                                const float2 l2 = bits_to_float2(lw[i]);
                                f[2 * i] += l2.x; f[2 * i + 1] += l2.y;
                            } else if (mode == 3) {
                                const float2 l2 = e4m3x2_to_float2((lw[i >> 1] >> (16 * (i & 1))) & 0xFFFFu);
                                f[2 * i] = fmaf(l2.x, 1.0f / LO_SCALE, f[2 * i]);
                                f[2 * i + 1] = fmaf(l2.y, 1.0f / LO_SCALE, f[2 * i + 1]);
                            }
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t c = (uint32_t)(2 * (2 * oc + e) + h) ^ px_sw;
                            *reinterpret_cast<float4*>(px_row + 4 * c) = make_float4(f[4 * h], f[4 * h + 1], f[4 * h + 2], f[4 * h + 3]);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < CHUNK_PX; ++j)
                        r[j] = __uint_as_float(stage[j * 32 + 4 * (lane_chunk ^ (uint32_t)(j & 7)) + lane_word]);
                    __syncwarp();
                    if (ch + 1 < ch_end) load_residual(ch + 1);    // in flight during this chunk's math and stores
                }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = vn[j];
                if (ch + 1 < ch_end) tmem_ld16_issue(taddr0 + (ch + 1) * CHUNK_PX, vn);
                if (!(p.debug & 1)) {
                    // ---- lane side: bias -> activation -> (BN affine * scale) -> + residual.  Padded channels
                    //      (co >= cout) come out as exact zeros: zero weights, bias 0, shift 0.
                    float f[16];
                    if (act == PCNN_ACT_LEAKY_RELU && !affine) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float a = fmaf(__uint_as_float(v[j]), asc, bias);
                            f[j] = fmaxf(a, 0.2f * a);
                        }
                    } else if (act == PCNN_ACT_LEAKY_RELU) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float a = fmaf(__uint_as_float(v[j]), asc, bias);
                            a = fmaxf(a, 0.2f * a);
                            f[j] = fmaf(a, mul, add);
                        }
                    } else if (act == PCNN_ACT_TANH && mode >= 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaf(tanh_fast(fmaf(__uint_as_float(v[j]), asc, bias)), mul, add);
                    } else if (act == PCNN_ACT_TANH) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaf(tanh_approx(fmaf(__uint_as_float(v[j]), asc, bias)), mul, add);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaf(fmaf(__uint_as_float(v[j]), asc, bias), mul, add);
                    }
                    if (has_res) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] += r[j];
                    }
                    // ---- rounding: one word per pixel with every output form, into the transpose buffer
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const __half2 hh = __floats2half2_rn(f[j], f[j + 1]);
                        const uint32_t hb = h2_bits(hh);
                        uint32_t w0, w1;
                        if (mode == 1) {
                            w0 = hb & 0xFFFFu; w1 = hb >> 16;
                        } else {
                            const float2 back = __half22float2(hh);
                            const float l0 = f[j] - back.x, l1 = f[j + 1] - back.y;
                            if (mode == 2) {
                                const uint32_t lb = h2_bits(__floats2half2_rn(l0, l1));
                                w0 = __byte_perm(hb, lb, 0x5410); w1 = __byte_perm(hb, lb, 0x7632);
                            } else {
                                const uint32_t qq = float2_to_e4m3x2(f[j], f[j + 1]) | (float2_to_e4m3x2(l0 * LO_SCALE, l1 * LO_SCALE) << 16);
                                w0 = __byte_perm(hb, qq, 0x6410); w1 = __byte_perm(hb, qq, 0x7532);
                            }
                        }
                        stage[j * 32 + 4 * (lane_chunk ^ (uint32_t)(j & 7)) + lane_word] = w0;
                        stage[(j + 1) * 32 + 4 * (lane_chunk ^ (uint32_t)((j + 1) & 7)) + lane_word] = w1;
                    }
                    __syncwarp();
                    // ---- pixel side: gather this pixel's two octets and store 16-byte units
                    const int xo = ch * CHUNK_PX;
                    const int x = x0 + xo + px;
                    const bool xok = x < p.W;
                    uint4 hi[2], lo[2];
                    uint2 q8[2], l8[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const uint4 a = *reinterpret_cast<const uint4*>(px_row + 4 * ((uint32_t)(2 * (2 * oc + e)) ^ px_sw));
                        const uint4 c = *reinterpret_cast<const uint4*>(px_row + 4 * ((uint32_t)(2 * (2 * oc + e) + 1) ^ px_sw));
                        hi[e] = make_uint4(__byte_perm(a.x, a.y, 0x5410), __byte_perm(a.z, a.w, 0x5410),
                                           __byte_perm(c.x, c.y, 0x5410), __byte_perm(c.z, c.w, 0x5410));
                        if (mode == 2) {
                            lo[e] = make_uint4(__byte_perm(a.x, a.y, 0x7632), __byte_perm(a.z, a.w, 0x7632),
                                               __byte_perm(c.x, c.y, 0x7632), __byte_perm(c.z, c.w, 0x7632));
                        } else if (mode == 3) {
                            const uint32_t t01 = __byte_perm(a.x, a.y, 0x7362), t23 = __byte_perm(a.z, a.w, 0x7362);
                            const uint32_t t45 = __byte_perm(c.x, c.y, 0x7362), t67 = __byte_perm(c.z, c.w, 0x7362);
                            q8[e] = make_uint2(__byte_perm(t01, t23, 0x5410), __byte_perm(t45, t67, 0x5410));
                            l8[e] = make_uint2(__byte_perm(t01, t23, 0x7632), __byte_perm(t45, t67, 0x7632));
                        }
                    }
                    // stores of both octets at a pixel displacement: (0,0) is the pixel itself, the others are its
                    // SYMMETRIC mirror images in the halo ring (fused tf.pad for the next layer; border pixels only)
                    auto emit = [&](bool ok0, bool ok1, long long d0, long long d1) {
                        const bool oks[2] = {ok0, ok1};
                        const long long ds[2] = {d0, d1};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (oks[e]) {
                                const size_t off = (size_t)((long long)o_off[e] + ((long long)xo + ds[e]) * 16);
                                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + off) = hi[e];
                                if (mode == 2) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out_lo) + off) = lo[e];
                            }
                        }
                        if (mode == 3) {
                            uint8_t* dq = reinterpret_cast<uint8_t*>(p.out_lo);
                            if (CP == 32 || CP == 16) {
                                // both octets belong to one row and one 16-channel group: planes (2g, 2g+1) -> 16-byte stores
                                if (ok0) {
                                    const size_t base = (size_t)((long long)o_off[0] + ((long long)xo + d0) * 16);
                                    if (o_pl[0] == tail_pl) {
                                        *reinterpret_cast<uint4*>(dq + base) = make_uint4(q8[0].x, q8[0].y, l8[0].x, l8[0].y);
                                    } else {
                                        *reinterpret_cast<uint4*>(dq + base) = make_uint4(q8[0].x, q8[0].y, q8[1].x, q8[1].y);
                                        *reinterpret_cast<uint4*>(dq + base + plane_px * 16) = make_uint4(l8[0].x, l8[0].y, l8[1].x, l8[1].y);
                                    }
                                }
                            } else {
                                // CP = 8 / 24: an octet is alone in its 16-channel group (the tail plane) or its neighbour lives in
                                // another thread (planes 0 / 1 of CP = 24: 8-byte stores)
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    if (oks[e]) {
                                        const size_t base = (size_t)((long long)o_off[e] + ((long long)xo + ds[e]) * 16);   // hi plane o_pl
                                        if (o_pl[e] == tail_pl) {            // alone in its group: [x | remainder] in its own q plane
                                            *reinterpret_cast<uint4*>(dq + base) = make_uint4(q8[e].x, q8[e].y, l8[e].x, l8[e].y);
                                        } else if (o_pl[e] == 0) {           // group 0, bytes 0..7
                                            *reinterpret_cast<uint2*>(dq + base) = q8[e];
                                            *reinterpret_cast<uint2*>(dq + base + plane_px * 16) = l8[e];
                                        } else {                             // plane 1: group 0, bytes 8..15 of planes 0 (x) and 1 (remainder)
                                            *reinterpret_cast<uint2*>(dq + base - plane_px * 16 + 8) = q8[e];
                                            *reinterpret_cast<uint2*>(dq + base + 8) = l8[e];
                                        }
                                    }
                                }
                            }
                        }
                    };
                    emit(o_ok[0] && xok, o_ok[1] && xok, 0, 0);
                    if (p.halo_sym && xok) {
                        // mirror displacements: rows -(2y+1) (top ring) / 2(H-y)-1 (bottom ring), columns likewise
                        const int dxm[3] = {0, (x < HALO) ? -(2 * x + 1) : 0, (x >= p.W - HALO) ? 2 * (p.W - x) - 1 : 0};
                        const bool any_y = (o_ok[0] && (o_dy[0][0] | o_dy[0][1])) || (o_ok[1] && (o_dy[1][0] | o_dy[1][1]));
                        if (any_y || dxm[1] || dxm[2]) {
#pragma unroll 1
                            for (int t = 1; t < 9; ++t) {
                                const int ai = t / 3, ci = t - 3 * ai;
                                if (ci && !dxm[ci]) continue;
                                const int dy0 = ai ? o_dy[0][ai - 1] : 0, dy1 = ai ? o_dy[1][ai - 1] : 0;
                                const bool k0 = o_ok[0] && (!ai || dy0), k1 = o_ok[1] && (!ai || dy1);
                                if (k0 || k1) emit(k0, k1, (long long)dy0 * p.P + dxm[ci], (long long)dy1 * p.P + dxm[ci]);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            if (lane == 0) mbar_arrive(acc_empty + acc);
        }
    }

    // teardown: everyone done with TMEM before it is released
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------- layout / packing kernels
// One 8-channel plane in the last 16-channel chunk (ceil(Cin/8) odd) and k >= 3: that chunk's fp16 passes pair column
// taps in the K dimension.  Packing and launching must agree, so both ask this function.  (PCNN_TC_NO_PAIR=1: A/B only.)
static inline bool pair_tail_for(int Cin, int k) {
    static const int off = getenv("PCNN_TC_NO_PAIR") ? atoi(getenv("PCNN_TC_NO_PAIR")) : 0;
    return !off && k >= 3 && (((Cin + 7) / 8) & 1);
}

// Keras kernel [kh][kw][Cin][Cout] fp32 -> packed fp16 [C16][kw][2][(kh+2(RT-1))*CP][8], RT = 128/CP.
// pair_tail: in the last chunk, stage s < (kw+1)/2 holds channels [16c, 16c+8) of column taps (2s, 2s+1) in its two K
// halves (last stage: a zero half and tap kw-1, read at column offset kw-2); the remaining stages of that chunk are unused.
__global__ void pack_weights_kernel(const float* __restrict__ k, __half* __restrict__ out, int kh, int kw,
                                    int Cin, int Cout, int c16, long long total, int nsplit, float scale, int cp, int pair_tail) {
    const int zpad = rows_per_tile(cp) - 1;
    const int Z = packed_zrows(kh, cp);
    const int npairs = (kw + 1) / 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int e = idx & 7;
        long long t = idx >> 3;
        const int co = t % cp; t /= cp;
        const int z = t % Z; t /= Z;
        const int pl = t & 1; t >>= 1;
        const int dx = t % kw;
        const int c = t / kw;
        const int dy = z - zpad;
        const bool rowok = dy >= 0 && dy < kh && co < Cout;
        auto w_at = [&](int tap, int ci) -> float {
            return (rowok && tap >= 0 && tap < kw && ci < Cin) ? scale * k[(((long long)dy * kw + tap) * Cin + ci) * Cout + co] : 0.f;
        };
        const float v = w_at(dx, c * 16 + pl * 8 + e);        // unpaired image: K half = channel half
        float vh = v;                                         // what the fp16 image holds
        if (pair_tail && c == c16 - 1) {
            int tap = -1;
            if (dx < npairs - 1) tap = 2 * dx + pl;
            else if (dx == npairs - 1 && pl == 1) tap = kw - 1;
            vh = w_at(tap, c * 16 + e);
        }
        const __half h = __float2half_rn(vh);
        out[idx] = h;
        if (nsplit == 2) out[total + idx] = __float2half_rn(vh - __half2float(h));
        if (nsplit == 3) {
            // fp8 image [c][dx][plane][z][co][16]: plane 0 = e4m3(W_lo) pairs with e4m3(x),
            // plane 1 = e4m3(W * 2^-11) pairs with e4m3(x_lo * 2^11); ci16 = pl*8+e of the fp16 image
            uint8_t* q = reinterpret_cast<uint8_t*>(out + total);
            const long long stage_elems = 2LL * Z * cp * 8;                    // fp16 elements per (c,dx) stage == bytes / 2
            const long long base = ((long long)c * kw + dx) * stage_elems * 2; // byte offset of the (c,dx) stage
            if (pair_tail && c == c16 - 1) {
                // paired tail chunk: K half pl = column tap (as in the fp16 image); its 16 bytes are [W_lo x 8 | W * 2^-11 x 8],
                // facing [e4m3(x) x 8 | e4m3(x_lo * 2^11) x 8] of the lone q plane
                const long long o = base + (((long long)pl * Z + z) * cp + co) * 16 + e;
                q[o] = to_e4m3(vh - __half2float(h));
                q[o + 8] = to_e4m3(vh * (1.0f / LO_SCALE));
            } else {
                const int ci16 = pl * 8 + e;
                q[base + (((long long)0 * Z + z) * cp + co) * 16 + ci16] = to_e4m3(v - __half2float(__float2half_rn(v)));
                q[base + (((long long)1 * Z + z) * cp + co) * 16 + ci16] = to_e4m3(v * (1.0f / LO_SCALE));
            }
        }
    }
}

// calls fn(pixel index within a padded plane) for interior pixel (y,x) and, with halo_sym, for its SYMMETRIC mirror
// images in the 7-wide halo ring (only border pixels have any)
template <class F>
__device__ __forceinline__ void mirror_targets(int y, int x, int H, int W, int P, int halo_sym, F&& fn) {
    fn((size_t)(y + HALO) * P + (x + HALO));
    if (!halo_sym) return;
    const int ys[3] = {y, (y < HALO) ? -1 - y : INT_MIN, (y >= H - HALO) ? 2 * H - 1 - y : INT_MIN};
    const int xs[3] = {x, (x < HALO) ? -1 - x : INT_MIN, (x >= W - HALO) ? 2 * W - 1 - x : INT_MIN};
#pragma unroll 1
    for (int t = 1; t < 9; ++t) {
        const int yy = ys[t / 3], xx = xs[t % 3];
        if (yy != INT_MIN && xx != INT_MIN) fn((size_t)(yy + HALO) * P + (xx + HALO));
    }
}

// NCHW fp32 -> BLK8 fp16 interior (planes [plane0, plane0 + ceil(C/8))).  grid (ceil(W/128), H, B*np): no
// per-element div/mod; each thread reads 8 channel planes (coalesced along x) and writes one 16-byte pixel.
__global__ void __launch_bounds__(128) to_blk8_kernel(const float* __restrict__ in, __half* __restrict__ out,
                                                      __half* __restrict__ out_lo, int C, int H, int W, int Hp, int P,
                                                      int c8_total, int plane0, long long in_bstride, int np, int halo_sym) {
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int b = blockIdx.z / np, pl = blockIdx.z - b * np;
    __align__(16) __half h[8];
    __align__(16) __half l[8];
    const float* src = in + (long long)b * in_bstride + ((long long)(pl * 8) * H + y) * W + x;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float f = (pl * 8 + e < C) ? __ldg(src + (long long)e * H * W) : 0.f;
        h[e] = __float2half_rn(f);
        l[e] = __float2half_rn(f - __half2float(h[e]));
    }
    const size_t plane = ((size_t)b * c8_total + plane0 + pl) * Hp * P;
    const uint4 hv = *reinterpret_cast<const uint4*>(h), lv = *reinterpret_cast<const uint4*>(l);
    mirror_targets(y, x, H, W, P, halo_sym, [&](size_t pix) {
        *reinterpret_cast<uint4*>(out + (plane + pix) * 8) = hv;
        if (out_lo) *reinterpret_cast<uint4*>(out_lo + (plane + pix) * 8) = lv;
    });
}

// mode 3 companion of to_blk8: fp8 planes 2c = e4m3(x), 2c+1 = e4m3((x - fp16(x)) * 2^11), 16 channels each
// tail_pl: plane of the tensor that is alone in its 16-channel group (odd live-plane count), or -1: that plane holds
// [e4m3(x) x 8 | e4m3(remainder) x 8] per pixel and its partner plane stays zero
__global__ void __launch_bounds__(128) to_q8_kernel(const float* __restrict__ in, uint8_t* __restrict__ outq, int C, int H,
                                                    int W, int Hp, int P, int c8_total, int plane0, long long in_bstride, int nc, int halo_sym,
                                                    int tail_pl) {
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int b = blockIdx.z / nc, cc = blockIdx.z - b * nc;
    __align__(16) uint8_t a8[16];
    __align__(16) uint8_t l8[16];
    const float* src = in + (long long)b * in_bstride + ((long long)(cc * 16) * H + y) * W + x;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const float f = (cc * 16 + e < C) ? __ldg(src + (long long)e * H * W) : 0.f;
        a8[e] = to_e4m3(f);
        l8[e] = to_e4m3((f - __half2float(__float2half_rn(f))) * LO_SCALE);
    }
    const size_t plane = ((size_t)b * c8_total + plane0 + 2 * cc) * Hp * P;
    const uint4 av = *reinterpret_cast<const uint4*>(a8), lv = *reinterpret_cast<const uint4*>(l8);
    const bool tail = (plane0 + 2 * cc == tail_pl);
    mirror_targets(y, x, H, W, P, halo_sym, [&](size_t pix) {
        if (tail) {
            *reinterpret_cast<uint4*>(outq + (plane + pix) * 16) = make_uint4(av.x, av.y, lv.x, lv.y);
        } else {
            *reinterpret_cast<uint4*>(outq + (plane + pix) * 16) = av;
            *reinterpret_cast<uint4*>(outq + (plane + pix + (size_t)Hp * P) * 16) = lv;
        }
    });
}

// BLK8 fp16 (+ remainder buffer) -> NCHW fp32.  grid (ceil(W/128), H, B*np): one 16-byte pixel read per thread,
// eight coalesced channel-plane writes.
__global__ void __launch_bounds__(128) from_blk8_kernel(const __half* __restrict__ in, const __half* __restrict__ in_lo,
                                                        float* __restrict__ out, int C, int H, int W, int Hp, int P,
                                                        int c8_total, int plane0, long long out_bstride, int np, int mode, int tail_pl) {
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int b = blockIdx.z / np, pl = blockIdx.z - b * np;
    const size_t off = ((((size_t)b * c8_total + plane0 + pl) * Hp + (y + HALO)) * P + (x + HALO)) * 8;
    const uint4 hv = *reinterpret_cast<const uint4*>(in + off);
    const __half* h = reinterpret_cast<const __half*>(&hv);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __half2float(h[e]);
    if (in_lo && mode == 2) {
        const uint4 lv = *reinterpret_cast<const uint4*>(in_lo + off);
        const __half* l = reinterpret_cast<const __half*>(&lv);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += __half2float(l[e]);
    }
    if (in_lo && mode == 3) {
        // channel c lives in fp8 plane 2*(c/16)+1 at byte c%16: this 8-channel plane is half of one 16-byte pixel
        // (the tail plane keeps its remainders in its own bytes 8..15)
        const int pa = plane0 + pl;
        const int qpl = (pa == tail_pl) ? pa : (pa | 1);
        const size_t qoff = ((((size_t)b * c8_total + qpl) * Hp + (y + HALO)) * P + (x + HALO)) * 16 + ((pa == tail_pl) ? 8 : 8 * (pa & 1));
        const uint2 qv = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(in_lo) + qoff);
        const uint8_t* q = reinterpret_cast<const uint8_t*>(&qv);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += from_e4m3(q[e]) * (1.0f / LO_SCALE);
    }
    float* dst = out + (long long)b * out_bstride + ((long long)(pl * 8) * H + y) * W + x;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (pl * 8 + e < C) dst[(long long)e * H * W] = f[e];
}

// halo fill of a BLK8 buffer: pad-wide ring around the interior, zero or SYMMETRIC mirror.
// grid (1, H + 2*pad, B*planes): interior rows only touch their 2*pad edge pixels.
__global__ void __launch_bounds__(128) blk8_halo_fill_kernel(__half* __restrict__ buf, int H, int W, int Hp, int P, int pad, int mode) {
    const int y = (int)blockIdx.y - pad;
    const size_t plane = (size_t)blockIdx.z * Hp * P;
    const bool edge_row = (y < 0) || (y >= H);
    const int n = edge_row ? W + 2 * pad : 2 * pad;
    const int sy = (mode == PCNN_PAD_SYMMETRIC) ? pad_src_index(y, H, PCNN_PAD_SYMMETRIC) : 0;
    for (int i = threadIdx.x; i < n; i += 128) {
        const int x = edge_row ? i - pad : (i < pad ? i - pad : W + (i - pad));
        uint4 v = make_uint4(0, 0, 0, 0);
        if (mode == PCNN_PAD_SYMMETRIC)
            v = *reinterpret_cast<const uint4*>(buf + (plane + (size_t)(sy + HALO) * P + (pad_src_index(x, W, PCNN_PAD_SYMMETRIC) + HALO)) * 8);
        *reinterpret_cast<uint4*>(buf + (plane + (size_t)(y + HALO) * P + (x + HALO)) * 8) = v;
    }
}

// DBCNN mode expansion straight into BLK8: out[b, m, x, y] = h[b,m,y] * S[m,x] * w[b,m]; channels M, M+1 = pos.
// grid (ceil(n/128), ceil(xres/EXP_ROWS), B): a thread owns one column y, keeps h[b,.,y] in registers and walks
// EXP_ROWS rows x (S[.,x] and w[b,.] staged in shared memory), writing 16-byte units: h is read once per 16 rows
// instead of once per row.
constexpr int EXP_ROWS = 16;
__global__ void __launch_bounds__(128) dbcnn_expand_blk8_kernel(const float* __restrict__ h, const float* __restrict__ S,
                                                                const float* __restrict__ mw, const float* __restrict__ posx,
                                                                const float* __restrict__ posy, __half* __restrict__ out,
                                                                __half* __restrict__ out_lo, int M, int xres, int n,
                                                                int c8_total, int np, int mode) {
    __shared__ float s_S[EXP_ROWS][32], s_w[32], s_px[EXP_ROWS];
    const int y = blockIdx.x * 128 + threadIdx.x, x0 = blockIdx.y * EXP_ROWS, b = blockIdx.z;
    for (int e = threadIdx.x; e < EXP_ROWS * 32; e += 128) {
        const int r = e >> 5, m = e & 31;
        s_S[r][m] = (m < M && x0 + r < xres) ? __ldg(S + (long long)m * xres + x0 + r) : 0.f;
    }
    if (threadIdx.x < 32) s_w[threadIdx.x] = threadIdx.x < M ? __ldg(mw + (long long)b * M + threadIdx.x) : 0.f;
    if (threadIdx.x < EXP_ROWS) s_px[threadIdx.x] = x0 + threadIdx.x < xres ? __ldg(posx + x0 + threadIdx.x) : 0.f;
    __syncthreads();
    if (y >= n) return;
    const int Hp = xres + 2 * HALO, P = n + 2 * HALO;
    const float py = __ldg(posy + y);
    float hv[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) hv[m] = m < M ? __ldg(h + ((long long)b * M + m) * n + y) : 0.f;
    const size_t plane_px = (size_t)Hp * P;
    for (int r = 0; r < EXP_ROWS && x0 + r < xres; ++r) {
        const size_t pix = (size_t)(x0 + r + HALO) * P + (y + HALO);
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
            if (pl < np) {
                __align__(16) __half v[8];
                __align__(16) __half l[8];
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int m = pl * 8 + e;
                    f[e] = hv[m] * s_S[r][m] * s_w[m];           // zero for m >= M
                    if (m == M) f[e] = s_px[r];
                    else if (m == M + 1) f[e] = py;
                    v[e] = __float2half_rn(f[e]);
                    l[e] = __float2half_rn(f[e] - __half2float(v[e]));
                }
                const size_t off = (((size_t)b * c8_total + pl) * plane_px + pix) * 8;
                *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(v);
                if (out_lo && mode == 2) *reinterpret_cast<uint4*>(out_lo + off) = *reinterpret_cast<const uint4*>(l);
                if (out_lo && mode == 3) {
                    // fp8 planes 2(pl/2), 2(pl/2)+1, bytes [8*(pl&1), +8): this plane is one half of its 16-channel group
                    uint8_t* q = reinterpret_cast<uint8_t*>(out_lo);
                    __align__(8) uint8_t a8[8];
                    __align__(8) uint8_t l8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        a8[e] = to_e4m3(f[e]);
                        l8[e] = to_e4m3((f[e] - __half2float(v[e])) * LO_SCALE);
                    }
                    // (a lone last plane keeps [x | remainder] in its own q plane)
                    const bool tail = (np & 1) && pl == np - 1;
                    const size_t q0 = (((size_t)b * c8_total + 2 * (pl >> 1)) * plane_px + pix) * 16 + 8 * (pl & 1);
                    const size_t q1 = tail ? q0 + 8 : q0 + plane_px * 16;
                    *reinterpret_cast<uint2*>(q + q0) = *reinterpret_cast<const uint2*>(a8);
                    *reinterpret_cast<uint2*>(q + q1) = *reinterpret_cast<const uint2*>(l8);
                }
            }
        }
    }
}

// The DBCNN's separable first layer (pcnn_conv2d_tc_rowweights) reads the boundary features as ONE row: BLK8 tensor
// [B][np][1+14][n+14][8], row HALO: channel m < M = h[b,m,y] * w[b,m]; channel M = 1 (pairs with posx[x] in the row weights),
// channel M+1 = posy[y].  One thread per (b, y) writes the np 16-byte units.
__global__ void __launch_bounds__(128) dbcnn_signal_blk8_kernel(const float* __restrict__ h, const float* __restrict__ mw,
                                                                const float* __restrict__ posy, __half* __restrict__ out,
                                                                int M, int n, int np) {
    const int y = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
    if (y >= n) return;
    const int P = n + 2 * HALO;
    const size_t plane_px = (size_t)(1 + 2 * HALO) * P;
    const size_t pix = (size_t)HALO * P + (y + HALO);
    for (int pl = 0; pl < np; ++pl) {
        __align__(16) __half v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int m = pl * 8 + e;
            float f = 0.f;
            if (m < M) f = __ldg(h + ((long long)b * M + m) * n + y) * __ldg(mw + (long long)b * M + m);
            else if (m == M) f = 1.f;
            else if (m == M + 1) f = __ldg(posy + y);
            v[e] = __float2half_rn(f);
        }
        *reinterpret_cast<uint4*>(out + (((size_t)b * (2 * ((np + 1) / 2)) + pl) * plane_px + pix) * 8) = *reinterpret_cast<const uint4*>(v);
    }
}

static inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
    long long g = (total + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// In-place  x += alpha * sum_r resize_r(src_r)  on channels [plane0*8, plane0*8 + 8*np) of a BLK8 tensor (all precision
// modes): the Upsample branches of the HPNN merge (layers/Upsample.py:56-59) for grids whose low-resolution sources are too
// large for the shared-memory staging of the fused upsample-merge kernel (2048^2: 64 x 64 x 32 floats per sample).
// One thread per (pixel, 8-channel plane): decode (as from_blk8), gather <= 16 taps x 8 channels per branch from the
// L2-resident sources, encode (as to_blk8 / to_q8).  No tail plane: the destination has an even number of planes.
constexpr int RA_MAX = 8;
struct ResizeAddParams {
    const float* src[RA_MAX];
    const int32_t* iy[RA_MAX]; const float* wy[RA_MAX];
    const int32_t* ix[RA_MAX]; const float* wx[RA_MAX];
    int taps[RA_MAX], ih[RA_MAX], iw[RA_MAX];
    int n, C;
    float alpha;
};
__global__ void __launch_bounds__(128) resize_add_blk8_kernel(const ResizeAddParams p, __half* __restrict__ buf, uint8_t* __restrict__ lo,
                                                              int mode, int H, int W, int Hp, int P, int c8_total, int plane0, int np) {
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int b = blockIdx.z / np, pl = blockIdx.z - b * np;
    const int pa = plane0 + pl;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < p.n; ++r) {
        const int taps = p.taps[r], ih = p.ih[r], iw = p.iw[r];
        const float* src = p.src[r] + ((long long)b * p.C + pl * 8) * ih * iw;
        for (int a = 0; a < taps; ++a) {
            const int sy = __ldg(p.iy[r] + y * taps + a);
            const float fy = __ldg(p.wy[r] + y * taps + a);
            for (int c = 0; c < taps; ++c) {
                const float w = fy * __ldg(p.wx[r] + x * taps + c);
                const float* s0 = src + (long long)sy * iw + __ldg(p.ix[r] + x * taps + c);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, __ldg(s0 + (long long)e * ih * iw), acc[e]);
            }
        }
    }
    const size_t pix = (size_t)(y + HALO) * P + (x + HALO);
    const size_t off = (((size_t)b * c8_total + pa) * Hp * P + pix) * 8;
    const uint4 hv = *reinterpret_cast<const uint4*>(buf + off);
    const __half* h = reinterpret_cast<const __half*>(&hv);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __half2float(h[e]);
    const size_t q0 = (((size_t)b * c8_total + (pa & ~1)) * Hp * P + pix) * 16 + 8 * (pa & 1);      // e4m3(x) bytes of this plane
    const size_t q1 = q0 + (size_t)Hp * P * 16;                                                     // remainder bytes
    if (mode == 2) {
        const uint4 lv = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(lo) + off);
        const __half* l = reinterpret_cast<const __half*>(&lv);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += __half2float(l[e]);
    } else if (mode == 3) {
        const uint2 qv = *reinterpret_cast<const uint2*>(lo + q1);
        const uint8_t* q = reinterpret_cast<const uint8_t*>(&qv);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += from_e4m3(q[e]) * (1.0f / LO_SCALE);
    }
    __align__(16) __half ho[8];
    __align__(16) __half lo16[8];
    __align__(8) uint8_t a8[8];
    __align__(8) uint8_t l8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float v = fmaf(p.alpha, acc[e], f[e]);
        ho[e] = __float2half_rn(v);
        const float rem = v - __half2float(ho[e]);
        lo16[e] = __float2half_rn(rem);
        a8[e] = to_e4m3(v);
        l8[e] = to_e4m3(rem * LO_SCALE);
    }
    *reinterpret_cast<uint4*>(buf + off) = *reinterpret_cast<const uint4*>(ho);
    if (mode == 2) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(lo) + off) = *reinterpret_cast<const uint4*>(lo16);
    if (mode == 3) {
        *reinterpret_cast<uint2*>(lo + q0) = *reinterpret_cast<const uint2*>(a8);
        *reinterpret_cast<uint2*>(lo + q1) = *reinterpret_cast<const uint2*>(l8);
    }
}

}  // namespace tc
}  // namespace pcnn

using namespace pcnn;
using namespace pcnn::tc;

// Mode-3 tensors with an odd number of live 8-channel planes: index of the last one (alone in its 16-channel group), else -1.
// Its q plane holds [e4m3(x) x 8 | e4m3((x - hi) * 2^11) x 8] per pixel and the partner plane stays zero.
static inline int tail_plane(int C) { const int np = (C + 7) / 8; return (np & 1) ? np - 1 : -1; }

extern "C" size_t pcnn_blk8_bytes(int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    const size_t planes = (size_t)((C + 15) / 16) * 2;
    // + slack: the last tile's row window may run a few hundred pixels past the final row
    return ((size_t)B * planes * (H + 2 * HALO) * (W + 2 * HALO) * 8 + 8192) * sizeof(__half);
}

// shared-memory budget of the kernel
constexpr size_t SMEM_MAX = 227 * 1024;
// large kernels are MMA-bound by a wide margin (4 epilogue warps leave shared memory for operands); k <= 5 layers are bound
// by the latency chain of their epilogue with two warps per scheduler, so they get three (PCNN_TC_EPI12=0: A/B only)
static inline int epi_warps_for(int k) {
    static const int epi12 = getenv("PCNN_TC_EPI12") ? atoi(getenv("PCNN_TC_EPI12")) : 1;
    return (k >= 11) ? 4 : ((k <= 5 && epi12) ? 12 : MAX_EPI_WARPS);
}
static inline size_t smem_fixed_for(int k) { return (size_t)epi_warps_for(k) * STAGE_WARP + 2048; }   // transpose buffers + mbarriers

// Channel slots per output row of the M operand for a k x k layer with Cout output channels: the narrowest of
// {8, 16, 32} >= Cout whose kh+RT-1 input rows (at the full 256-pixel tile width) plus three weight stages fit in
// shared memory.  Depends only on (Cout, k), so packing and launching always agree.
// (PCNN_TC_CP=32 forces the widest layout: A/B experiments only.)
static int choose_cp(int cout, int k) {
    static const int forced = getenv("PCNN_TC_CP") ? atoi(getenv("PCNN_TC_CP")) : 0;
    if (forced == 32) return 32;
    static const int no24 = getenv("PCNN_TC_NO_CP24") ? atoi(getenv("PCNN_TC_NO_CP24")) : 0;
    for (int cp = 8; cp < 32; cp += 8) {
        if (cp < cout || (cp == 24 && no24)) continue;
        const int zpad = rows_per_tile(cp) - 1, R = k + zpad;
        const size_t rowslot = 2 * (((size_t)(256 + k - 1) * 16 + 127) & ~(size_t)127);
        const size_t wst = 2 * (size_t)packed_zrows(k, cp) * cp * 16;
        if ((size_t)R * rowslot + 3 * wst + smem_fixed_for(k) <= SMEM_MAX) return cp;
    }
    return 32;
}

extern "C" int pcnn_conv_tc_channel_slots(int Cout, int k) { return (Cout >= 1 && Cout <= 32) ? choose_cp(Cout, k) : 0; }

extern "C" size_t pcnn_conv_tc_packed_weight_bytes(int kh, int kw, int Cin, int Cout, int nsplit) {
    if (Cout < 1 || Cout > 32) return 0;
    const int cp = choose_cp(Cout, kh);
    return (size_t)(nsplit >= 2 ? 2 : 1) * ((Cin + 15) / 16) * kw * 2 * packed_zrows(kh, cp) * cp * 8 * sizeof(__half);
}

extern "C" int pcnn_conv_tc_pack_weights(const float* kernel, void* packed, int kh, int kw, int Cin, int Cout, int nsplit,
                                         float scale, void* stream) {
    PCNN_CHECK_ARG(scale > 0.f, "conv_tc_pack_weights: scale must be positive (a power of two)");
    PCNN_CHECK_ARG(nsplit >= 1 && nsplit <= 3, "conv_tc_pack_weights: precision mode must be 1, 2 or 3");
    PCNN_CHECK_ARG(kernel && packed, "conv_tc_pack_weights: null pointer");
    PCNN_CHECK_ARG(kh == kw && (kh & 1) && kh >= 1 && kh <= 2 * HALO + 1, "conv_tc: kernel %dx%d not supported (odd, square, <= 15)", kh, kw);
    PCNN_CHECK_ARG(Cout >= 1 && Cout <= 32 && Cin >= 1, "conv_tc: Cout %d not in [1,32]", Cout);
    const int c16 = (Cin + 15) / 16, cp = choose_cp(Cout, kh);
    const long long total = (long long)c16 * kw * 2 * packed_zrows(kh, cp) * cp * 8;
    pack_weights_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(kernel, (__half*)packed, kh, kw, Cin, Cout, c16, total, nsplit, scale, cp,
                                                                           pair_tail_for(Cin, kh) ? 1 : 0);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_to_blk8(const float* in, void* out, void* out_lo, int mode, int B, int C, int H, int W, int c_total, int c_offset,
                            int64_t in_bstride, int halo_mode, void* stream) {
    PCNN_CHECK_ARG(halo_mode == PCNN_PAD_CONSTANT || (halo_mode == PCNN_PAD_SYMMETRIC && H >= HALO && W >= HALO),
                   "to_blk8: halo_mode must be 0 (halo untouched) or SYMMETRIC on a map of at least 7x7");
    const int hsym = halo_mode == PCNN_PAD_SYMMETRIC;
    PCNN_CHECK_ARG(mode >= 1 && mode <= 3 && (mode == 1 || out_lo), "to_blk8: precision mode 2/3 needs the second buffer");
    PCNN_CHECK_ARG(mode != 3 || (c_offset % 16) == 0, "to_blk8: mode 3 needs a channel offset that is a multiple of 16");
    PCNN_CHECK_ARG(in && out && B > 0 && C > 0 && (c_offset % 8) == 0 && c_offset + C <= ((c_total + 15) / 16) * 16, "to_blk8: bad argument");
    const int tail_pl = tail_plane(c_total);
    const int c8_total = ((c_total + 15) / 16) * 2;
    const int np = (C + 7) / 8, nc = (C + 15) / 16;
    PCNN_CHECK_ARG(H <= 65535 && (long long)B * np <= 65535, "to_blk8: grid too large (H %d, B*planes %lld)", H, (long long)B * np);
    to_blk8_kernel<<<dim3(ceil_div(W, 128), H, B * np), 128, 0, (cudaStream_t)stream>>>(in, (__half*)out, mode == 2 ? (__half*)out_lo : nullptr, C, H, W, H + 2 * HALO, W + 2 * HALO, c8_total, c_offset / 8, in_bstride, np, hsym);
    PCNN_CHECK_LAUNCH();
    if (mode == 3) {
        to_q8_kernel<<<dim3(ceil_div(W, 128), H, B * nc), 128, 0, (cudaStream_t)stream>>>(in, (uint8_t*)out_lo, C, H, W, H + 2 * HALO, W + 2 * HALO, c8_total, c_offset / 8, in_bstride, nc, hsym, tail_pl);
        PCNN_CHECK_LAUNCH();
    }
    return PCNN_OK;
}

extern "C" int pcnn_from_blk8(const void* in, const void* in_lo, int mode, float* out, int B, int C, int H, int W, int c_total, int c_offset,
                              int64_t out_bstride, void* stream) {
    PCNN_CHECK_ARG(mode >= 1 && mode <= 3, "from_blk8: bad precision mode");
    PCNN_CHECK_ARG(mode != 3 || (c_offset % 16) == 0, "from_blk8: mode 3 needs a channel offset that is a multiple of 16");
    PCNN_CHECK_ARG(in && out && B > 0 && C > 0 && (c_offset % 8) == 0, "from_blk8: bad argument");
    const int c8_total = ((c_total + 15) / 16) * 2;
    const int np = (C + 7) / 8;
    PCNN_CHECK_ARG(H <= 65535 && (long long)B * np <= 65535, "from_blk8: grid too large");
    from_blk8_kernel<<<dim3(ceil_div(W, 128), H, B * np), 128, 0, (cudaStream_t)stream>>>((const __half*)in, mode >= 2 ? (const __half*)in_lo : nullptr, out, C, H, W, H + 2 * HALO, W + 2 * HALO, c8_total, c_offset / 8, out_bstride, np, mode, tail_plane(c_total));
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_blk8_halo_fill(void* buf, int B, int C, int H, int W, int pad, int mode, void* stream) {
    PCNN_CHECK_ARG(buf && B > 0 && C > 0 && pad >= 0 && pad <= HALO, "blk8_halo_fill: bad argument");
    PCNN_CHECK_ARG(mode == PCNN_PAD_CONSTANT || mode == PCNN_PAD_SYMMETRIC, "blk8_halo_fill: mode must be CONSTANT(0) or SYMMETRIC");
    if (mode == PCNN_PAD_SYMMETRIC) PCNN_CHECK_ARG(pad <= H && pad <= W, "blk8_halo_fill: SYMMETRIC pad %d larger than the tensor (%d,%d)", pad, H, W);
    if (pad == 0) return PCNN_OK;
    const int planes = ((C + 15) / 16) * 2;
    PCNN_CHECK_ARG(H + 2 * pad <= 65535 && (long long)B * planes <= 65535, "blk8_halo_fill: grid too large");
    blk8_halo_fill_kernel<<<dim3(1, H + 2 * pad, B * planes), 128, 0, (cudaStream_t)stream>>>((__half*)buf, H, W, H + 2 * HALO, W + 2 * HALO, pad, mode);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_dbcnn_expand_blk8(const float* h, const float* sinh_basis, const float* modew, const float* posx,
                                      const float* posy, void* out, void* out_lo, int mode, int B, int M, int xres, int n, void* stream) {
    PCNN_CHECK_ARG(h && sinh_basis && modew && posx && posy && out && B > 0 && M > 0, "dbcnn_expand_blk8: bad argument");
    const int c8_total = ((M + 2 + 15) / 16) * 2;
    const int np = (M + 2 + 7) / 8;
    PCNN_CHECK_ARG(xres <= 65535 && B <= 65535 && M + 2 <= 32, "dbcnn_expand_blk8: grid too large or more than 30 modes");
    dbcnn_expand_blk8_kernel<<<dim3(ceil_div(n, 128), ceil_div(xres, EXP_ROWS), B), 128, 0, (cudaStream_t)stream>>>(h, sinh_basis, modew, posx, posy, (__half*)out, (__half*)out_lo, M, xres, n, c8_total, np, mode);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

static int conv2d_tc_impl(const void* in, const void* in_lo, const void* wpack, const float* bias, const float* bn_scale,
                          const float* bn_shift, const void* residual, const void* residual_lo, const float* out_scale,
                          void* out, void* out_lo, int B, int Cin_total, int Cout, int Cout_total, int Cres_total,
                          int H, int W, int k, int act, int nsplit, float acc_scale, int out_halo_mode, int num_sms, void* stream,
                          int rw, int rw_tslots) {
    PCNN_CHECK_ARG(in && wpack && out, "conv2d_tc: null pointer");
    const bool skip_corr = (nsplit & PCNN_TC_SKIP_CORRECTION) != 0;
    nsplit &= ~PCNN_TC_SKIP_CORRECTION;
    PCNN_CHECK_ARG(nsplit >= 1 && nsplit <= 3, "conv2d_tc: precision mode must be 1, 2 or 3");
    PCNN_CHECK_ARG(!skip_corr || nsplit == 3, "conv2d_tc: PCNN_TC_SKIP_CORRECTION applies to precision mode 3 only");
    if (nsplit >= 2) PCNN_CHECK_ARG(in_lo && out_lo && (!residual || residual_lo), "conv2d_tc: split precision needs the lo buffers");
    PCNN_CHECK_ARG((k & 1) && k >= 1 && k <= 2 * HALO + 1, "conv2d_tc: kernel size %d not supported (odd, <= 15)", k);
    PCNN_CHECK_ARG(Cout >= 1 && Cout <= 32 && Cout <= Cout_total, "conv2d_tc: Cout %d not in [1,32]", Cout);
    PCNN_CHECK_ARG(B > 0 && H > 0 && W > 0 && Cin_total > 0, "conv2d_tc: bad shape");
    PCNN_CHECK_ARG((bn_scale == nullptr) == (bn_shift == nullptr), "conv2d_tc: bn_scale/bn_shift must come together");
    PCNN_CHECK_ARG(out_halo_mode == PCNN_PAD_CONSTANT || (out_halo_mode == PCNN_PAD_SYMMETRIC && H >= HALO && W >= HALO),
                   "conv2d_tc: out_halo_mode must be 0 (halo untouched) or SYMMETRIC on a map of at least 7x7");
    Params p;
    p.in = (const __half*)in; p.in_lo = (const __half*)in_lo; p.wpack = (const __half*)wpack; p.bias = bias;
    p.bn_scale = bn_scale; p.bn_shift = bn_shift;
    p.residual = (const __half*)residual; p.residual_lo = (const __half*)residual_lo; p.out_scale = out_scale;
    p.out = (__half*)out; p.out_lo = (__half*)out_lo; p.nsplit = nsplit; p.acc_scale = acc_scale;
    p.B = B; p.H = H; p.W = W; p.Hp = H + 2 * HALO; p.P = W + 2 * HALO;
    p.c16 = (Cin_total + 15) / 16;
    p.kpass = skip_corr ? 1 : (nsplit == 2 ? 3 : (nsplit == 3 ? 2 : 1));
    p.nv = p.c16 * p.kpass;
    p.pair_tail = (!rw && pair_tail_for(Cin_total, k)) ? 1 : 0;
    p.npairs = (k + 1) / 2;
    p.rw = rw; p.rw_tslots = rw_tslots;
    p.res_tail_pl = (nsplit == 3 && residual) ? tail_plane(Cres_total) : -1;
    // mode 3: a lone last plane of the output uses the tail q layout, which belongs to the TENSOR: a convolution with an odd
    // plane count must own the tensor's last plane (and then the tensor has as many live planes as the convolution writes)
    PCNN_CHECK_ARG(nsplit != 3 || (((Cout + 7) / 8) & 1) == 0 || (Cout_total + 7) / 8 == (Cout + 7) / 8,
                   "conv2d_tc: precision mode 3 cannot write %d channels into a %d-channel tensor (odd plane count)", Cout, Cout_total);
    PCNN_CHECK_ARG(nsplit != 3 || (((Cout + 7) / 8) & 1) || tail_plane(Cout_total) < 0 || (Cout + 7) / 8 <= tail_plane(Cout_total),
                   "conv2d_tc: precision mode 3: %d channels overlap the tail plane of a %d-channel tensor", Cout, Cout_total);
    p.c8_in = p.c16 * 2; p.c8_out = ((Cout_total + 15) / 16) * 2; p.c8_res = ((Cres_total + 15) / 16) * 2;
    p.cout = Cout; p.kh = k; p.kw = k; p.pad = k / 2; p.act = act; p.halo_sym = (out_halo_mode == PCNN_PAD_SYMMETRIC);
    const int cp = choose_cp(Cout, k), rt = rows_per_tile(cp), zpad = rt - 1;
    p.n_tile = W >= 256 ? 256 : ((W + 15) / 16) * 16;
    p.tiles_x = ceil_div(W, p.n_tile); p.tiles_y = ceil_div(H, rt);
    p.num_tiles = B * p.tiles_x * p.tiles_y;
    p.row_copy_bytes = (uint32_t)(p.n_tile + k - 1) * 16u;
    p.rowplane_bytes = (p.row_copy_bytes + 127u) & ~127u;
    p.wstage_bytes = rw ? 4096u : 2u * (uint32_t)packed_zrows(k, cp) * (uint32_t)cp * 16u;
    // instruction descriptor: D=F32, A=B=F16, both K-major, N = n_tile, M = 128
    p.idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // shared-memory plan.  Rows need >= R slots (ideally 2R: full double buffering across chunk switches).  Weights:
    // if every stage of a tile (nv*kw of them) fits, they stay RESIDENT for the whole kernel; otherwise as many
    // stages as fit (<= 12) are kept in flight -- the weight stream is latency-bound with few small stages.
    const size_t kMax = SMEM_MAX;
    // (12 warps only for single-pass tensors: with 128 registers the mode-2/3 epilogues of the 480-thread build are no faster)
    p.n_epi = (epi_warps_for(k) == 12 && nsplit != 1) ? MAX_EPI_WARPS : epi_warps_for(k);
    if (rw) p.n_epi = 12;                         // 14 MMAs per tile: this layer is all epilogue
    const size_t fixed = rw ? (size_t)12 * STAGE_WARP + 2048      // shared-memory plan as for the widest epilogue: choose_cp() must not
                            : smem_fixed_for(k);                  // depend on the mode
    const int R = rw ? 1 : k + zpad;
    const size_t avail = kMax - fixed, rowslot = 2 * (size_t)p.rowplane_bytes, wst = p.wstage_bytes;
    PCNN_CHECK_ARG((size_t)R * rowslot + 2 * wst <= avail, "conv2d_tc: tile does not fit in shared memory (k=%d, n_tile=%d)", k, p.n_tile);
    // weight stages per tile: kw per pass, (kw+1)/2 for the passes of a paired last chunk
    const int total_stages = p.nv * k - (p.pair_tail ? p.kpass * (k - p.npairs) : 0);
    int slots, w_stages, resident = 0;
    if (rw) {
        slots = 4; w_stages = 12;                 // four tiles' signal rows and twelve 4 KB weight stages in flight
    } else if ((size_t)R * rowslot + (size_t)total_stages * wst <= avail && total_stages <= 48) {
        resident = 1; w_stages = total_stages;
        slots = (int)std::min<size_t>(2 * R, (avail - (size_t)w_stages * wst) / rowslot);
    } else {
        slots = 2 * R;
        w_stages = 0;
        while (slots >= R) {
            if ((size_t)slots * rowslot < avail) w_stages = (int)std::min<size_t>(12, (avail - (size_t)slots * rowslot) / wst);
            if (w_stages >= 6 || (slots == R && w_stages >= 2)) break;
            --slots;
        }
        PCNN_CHECK_ARG(slots >= R && w_stages >= 2, "conv2d_tc: shared-memory plan failed (k=%d, n_tile=%d)", k, p.n_tile);
    }
    if (R & 1) slots = (slots / R) * R;           // odd R: groups of (R+1)/2 and R/2 rows, the ring wraps between chunks only
    else slots = (slots / (R / 2)) * (R / 2);     // barrier groups of R/2 rows must not straddle the ring wrap
    PCNN_CHECK_ARG(2 * (slots + w_stages) + 5 <= 250, "conv2d_tc: too many pipeline stages for the mbarrier area");
    p.row_slots = slots; p.w_stages = w_stages; p.w_resident = resident;
    { static const int dbg = getenv("PCNN_TC_DEBUG") ? atoi(getenv("PCNN_TC_DEBUG")) : 0; p.debug = dbg; }
    const size_t smem = (size_t)slots * rowslot + (size_t)w_stages * wst + fixed;
    if (num_sms <= 0) num_sms = 148;
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    auto launch = [&](auto kern) -> int {
        PCNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMax));
        kern<<<grid, (3 + (p.n_epi > 8 ? 12 : 8)) * 32, smem, (cudaStream_t)stream>>>(p);
        PCNN_CHECK_LAUNCH();
        return PCNN_OK;
    };
#define PCNN_TC_DISPATCH(CPV)                                                                                   \
    if (cp == CPV && p.n_epi <= 8) {                                                                            \
        if (nsplit == 1) return launch(conv_tc_kernel<CPV, 1, 8>);                                              \
        if (nsplit == 2) return launch(conv_tc_kernel<CPV, 2, 8>);                                              \
        return launch(conv_tc_kernel<CPV, 3, 8>);                                                               \
    }                                                                                                           \
    if (cp == CPV) {                                                                                            \
        if (nsplit == 1) return launch(conv_tc_kernel<CPV, 1, 12>);                                             \
        if (nsplit == 2) return launch(conv_tc_kernel<CPV, 2, 12>);                                             \
        return launch(conv_tc_kernel<CPV, 3, 12>);                                                              \
    }
    PCNN_TC_DISPATCH(32)
    PCNN_TC_DISPATCH(24)
    PCNN_TC_DISPATCH(16)
    PCNN_TC_DISPATCH(8)
#undef PCNN_TC_DISPATCH
    return PCNN_ERR_INVALID_ARGUMENT;
}

extern "C" int pcnn_dbcnn_signal_blk8(const float* h, const float* modew, const float* posy, void* out, int B, int M, int n, void* stream) {
    PCNN_CHECK_ARG(h && modew && posy && out && B > 0 && B <= 65535 && M > 0 && M + 2 <= 32 && n > 0, "dbcnn_signal_blk8: bad argument");
    dbcnn_signal_blk8_kernel<<<dim3(ceil_div(n, 128), B), 128, 0, (cudaStream_t)stream>>>(h, modew, posy, (__half*)out, M, n, (M + 2 + 7) / 8);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_conv2d_tc(const void* in, const void* in_lo, const void* wpack, const float* bias, const float* bn_scale,
                              const float* bn_shift, const void* residual, const void* residual_lo, const float* out_scale,
                              void* out, void* out_lo, int B, int Cin_total, int Cout, int Cout_total, int Cres_total,
                              int H, int W, int k, int act, int nsplit, float acc_scale, int out_halo_mode, int num_sms, void* stream) {
    return conv2d_tc_impl(in, in_lo, wpack, bias, bn_scale, bn_shift, residual, residual_lo, out_scale, out, out_lo, B, Cin_total, Cout,
                          Cout_total, Cres_total, H, W, k, act, nsplit, acc_scale, out_halo_mode, num_sms, stream, 0, 0);
}

// Slots of the row-weight array for an H-row output: tiles*RT + 2 (the 128-row window of the last tile stays inside).
extern "C" int pcnn_conv_tc_rowweight_slots(int Cout, int k, int H) {
    if (Cout < 1 || Cout > 32 || H < 1) return 0;
    const int rt = rows_per_tile(choose_cp(Cout, k));
    return ceil_div(H, rt) * rt + 2;
}

extern "C" int pcnn_conv2d_tc_rowweights(const void* in_row, const void* wrow, const float* bias, void* out, int B, int Cin, int Cout,
                                         int Cout_total, int H, int W, int k, int act, float acc_scale, int num_sms, void* stream) {
    PCNN_CHECK_ARG(in_row && wrow && out, "conv2d_tc_rowweights: null pointer");
    return conv2d_tc_impl(in_row, nullptr, wrow, bias, nullptr, nullptr, nullptr, nullptr, nullptr, out, nullptr, B, Cin, Cout, Cout_total, 0,
                          H, W, k, act, 1, acc_scale, PCNN_PAD_CONSTANT, num_sms, stream, 1, pcnn_conv_tc_rowweight_slots(Cout, k, H));
}

extern "C" int pcnn_resize_add_blk8(int n_resize, const float* const* rs_in, const int32_t* const* rs_iy, const float* const* rs_wy,
                                    const int32_t* const* rs_ix, const float* const* rs_wx, const int* rs_taps, const int* rs_ih,
                                    const int* rs_iw, float alpha, void* buf, void* buf_lo, int mode, int B, int C, int H, int W,
                                    int c_total, int c_offset, void* stream) {
    PCNN_CHECK_ARG(n_resize >= 1 && n_resize <= RA_MAX && rs_in && rs_iy && rs_wy && rs_ix && rs_wx && rs_taps && rs_ih && rs_iw,
                   "resize_add_blk8: between 1 and %d resize branches", RA_MAX);
    PCNN_CHECK_ARG(buf && mode >= 1 && mode <= 3 && (mode == 1 || buf_lo), "resize_add_blk8: bad destination / precision mode");
    PCNN_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (c_offset % 16) == 0 && c_offset + C <= ((c_total + 15) / 16) * 16,
                   "resize_add_blk8: C must be a multiple of 8, the channel offset a multiple of 16 inside the buffer");
    PCNN_CHECK_ARG(mode != 3 || ((((c_total + 7) / 8) & 1) == 0 && (C % 16) == 0), "resize_add_blk8: precision mode 3 needs whole 16-channel groups and no tail plane");
    ResizeAddParams p;
    p.n = n_resize; p.C = C; p.alpha = alpha;
    for (int r = 0; r < n_resize; ++r) {
        PCNN_CHECK_ARG(rs_in[r] && rs_iy[r] && rs_wy[r] && rs_ix[r] && rs_wx[r] && rs_taps[r] >= 1 && rs_taps[r] <= 4 && rs_ih[r] > 0 && rs_iw[r] > 0,
                       "resize_add_blk8: branch %d: bad argument", r);
        p.src[r] = rs_in[r]; p.iy[r] = rs_iy[r]; p.wy[r] = rs_wy[r]; p.ix[r] = rs_ix[r]; p.wx[r] = rs_wx[r];
        p.taps[r] = rs_taps[r]; p.ih[r] = rs_ih[r]; p.iw[r] = rs_iw[r];
    }
    const int np = C / 8;
    PCNN_CHECK_ARG(H <= 65535 && (long long)B * np <= 65535, "resize_add_blk8: grid too large");
    resize_add_blk8_kernel<<<dim3(ceil_div(W, 128), H, B * np), 128, 0, (cudaStream_t)stream>>>(
        p, (__half*)buf, (uint8_t*)buf_lo, mode, H, W, H + 2 * HALO, W + 2 * HALO, ((c_total + 15) / 16) * 2, c_offset / 8, np);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
