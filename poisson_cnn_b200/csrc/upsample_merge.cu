// Fused upsample + merge of the HPNN bottleneck branches (models/Homogeneous_Poisson_NN_Legacy.py:226-233 of the
// reference sums the upsampled branch outputs; blocks/bottleneck_block.py:57-118 upsample each branch with
// deconvupscale (k == stride transpose conv, layers/deconvupscale.py:100-109) or tf.image.resize
// (layers/Upsample.py:57)).  Persistent CTAs produce output row segments of 256 pixels x 32 channels, one branch
// ("stage") at a time, the next stage's operands streaming into shared memory (cp.async) meanwhile:
//   * every deconv branch: the low-res row that feeds output row Y and the s phase matrices W[ty][*] sit in
//     shared memory; thread groups (column phase tx, 8 channels, 4 low-res pixels) compute all column phases
//     at once on the FP32 FMA pipe, stage them phase-major in shared memory, and each
//     of the first 256 threads -- which owns ONE output pixel, all channels -- gathers its phase
//     (conflict-free), applies bias + activation and adds it to its running sums (32 registers);
//   * every resize branch: the tiny source map sits in shared memory, each thread interpolates its pixel;
//   * the sum (x alpha) is written ONCE, straight into the tensor-core operand layout (BLK8 fp16 + remainder /
//     e4m3 planes) at a channel offset of the concat buffer: the full-resolution fp32 `merged` tensor and its
//     eight read-modify-write passes never exist.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <algorithm>

#include "pcnn_common.cuh"

#ifdef PCNN_UM_PROFILE
__device__ unsigned long long g_um_prof[16];
#define UM_T(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_um_prof[i], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)
#else
#define UM_T(i) do { } while (0)
#endif

namespace pcnn {
namespace um {

constexpr int HALO = 7;
constexpr int MAXB = 8;          // branches of each kind
constexpr int MAXS = 40;         // stages: (deconv branch, group of <= TXG column phases) or resize branch
constexpr int TXG = 8;           // column phases per stage: bounds the staged phase matrices to 8 x C x C floats
constexpr int SEG_W = 256;       // output pixels per row segment
constexpr int NTHR = 288;        // threads per CTA: 256 pixel owners + one warp (s = 3 has 264 work items).  Measured: 256-pixel
                                 // segments with one CTA per SM (4.05 ms per 64 samples) beat 128-pixel segments with two (4.4 ms):
                                 // the per-stage fixed cost is paid half as often
constexpr float LO_SCALE = 2048.f;

struct Params {
    int n_dc, n_rs;
    const float* dc_in[MAXB];    // [B][Cin][ih][iw]
    const float* dc_w[MAXB];     // packed deconv kernel [s][s][C/8][8*C + 4] (pcnn_upsample_merge_pack_kernel)
    const float* dc_b[MAXB];     // [Cout] or null
    int dc_s[MAXB], dc_ih[MAXB], dc_iw[MAXB], dc_pbh[MAXB], dc_pbw[MAXB], dc_act[MAXB], dc_ps[MAXB];
    const float* rs_in[MAXB];    // [B][C][ih][iw]
    const int* rs_iy[MAXB]; const float* rs_wy[MAXB]; const int* rs_ix[MAXB]; const float* rs_wx[MAXB];
    int rs_taps[MAXB], rs_ih[MAXB], rs_iw[MAXB];
    int rs_off[MAXB], rs_tab[MAXB], rs_st[MAXB];   // offsets: source / interpolation tables in the operand buffer, row buffer in s_st
    int dc_al16[MAXB];             // deconv input pointer is 16-byte aligned
    uint32_t dc_magic[MAXB];       // ceil(2^32 / stride)
    int n_stages;
    signed char st_branch[MAXS];   // deconv branch index, or -1 - r for resize branch r
    unsigned char st_tx0[MAXS], st_ntx[MAXS];
    float alpha;
    __half* out; uint8_t* out_lo;
    int mode;                    // precision mode of the destination (1, 2, 3: see pcnn_conv2d_tc)
    int B, C, H, W, c8_total, plane0;
    int tail_pl;             // mode 3: plane of the tensor that is alone in its 16-channel group (odd live-plane count), or -1
    int st_floats;               // staging region (phase-major deconv results / interpolated resize row)
    int tile_floats;             // offset of the weights inside an operand buffer (= size of the largest input tile)
    int buf_floats[2];           // operand buffers used by even / odd steps
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {
    const uint32_t lo = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 4 : 0;      // src-size 0: the 4 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {     // L2 -> shared, bypassing L1
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16z(float* dst, const float* src, bool valid) {   // zero-filled when !valid
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16p(float* dst, const float* src, int bytes) {   // first `bytes` copied, rest zero-filled
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(bytes > 0 ? src : nullptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_copy_g2s(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// geometry of deconv branch d for one row segment
struct DcGeom { int s, ih, iw, i, ty, j_begin, segp, pgs; };
// n / s for 0 <= n < 2^20 and s <= 64 through the branch's precomputed ceil(2^32 / s)
__device__ __forceinline__ int fast_div(int n, uint32_t magic) { return (int)__umulhi((uint32_t)n, magic); }
__device__ __forceinline__ DcGeom dc_geom(const Params& p, int d, int X0, int Y) {
    DcGeom g;
    g.s = p.dc_s[d]; g.ih = p.dc_ih[d]; g.iw = p.dc_iw[d];
    const uint32_t mg = p.dc_magic[d];
    const int jy = Y + p.dc_pbh[d];
    g.i = fast_div(jy, mg); g.ty = jy - g.i * g.s;
    const int xlast = min(X0 + SEG_W - 1, p.W - 1);
    g.j_begin = fast_div(X0 + p.dc_pbw[d], mg);
    const int j_end = fast_div(xlast + p.dc_pbw[d], mg);
    g.segp = ((j_end - g.j_begin + 1) + 3) & ~3; g.pgs = g.segp >> 2;
    return g;
}

// Persistent CTAs walk (row segment, stage) pairs; stage = one branch.  The operands of the NEXT stage
// (low-res row tile + the s phase matrices W[ty][0..s), or the tiny resize source) stream into the other
// shared-memory buffer with cp.async while the current stage computes, also across row boundaries.
__global__ void __launch_bounds__(NTHR, 1) upsample_merge_kernel(const Params p) {
    extern __shared__ __align__(16) float sm[];
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm);      // two mbarriers: bulk-copied phase matrices of each operand buffer
    float* s_st = sm + 4;                  // [s column phases][phase stride] staged deconv results / resize row
    float* s_buf[2] = {s_st + p.st_floats, s_st + p.st_floats + p.buf_floats[0]};   // stage operands: [tile | weights | bias] or [source | tables]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = p.C;
    const int nstages = p.n_stages;
    const int segs = (p.W + SEG_W - 1) / SEG_W;
    const int units = p.B * p.H * segs;
    const int my_units = (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_units * nstages;
    const int wblk = 8 * C + 4;            // staged phase matrices: (tx, channel group) blocks of 8 rows x C floats, 16 B apart in banks
    const int ncg = C >> 3;
    if (tid == 0) {
        mbar_init(s_bar, 1); mbar_init(s_bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto unit_coords = [&](int k, int& b, int& Y, int& X0) {
        const int u = (int)blockIdx.x + k * (int)gridDim.x;
        const int seg = u % segs;
        const int t = u / segs;
        Y = t % p.H; b = t / p.H; X0 = seg * SEG_W;
    };

    // asynchronous operand load of step (unit k, stage st): one cp.async group
    auto issue_load = [&](int step, int st, int b, int Y, int X0) {
        float* dst = s_buf[step & 1];
        const int br = p.st_branch[st];
        if (br >= 0) {
            const DcGeom g = dc_geom(p, br, X0, Y);
            const float* inb = p.dc_in[br] + ((long long)b * C * g.ih + g.i) * g.iw + g.j_begin;
            const long long cstride = (long long)g.ih * g.iw;
            constexpr int NW = NTHR / 32;
            if (((g.iw | g.j_begin) & 3) == 0 && p.dc_al16[br]) {
                // rows are 16-byte aligned: whole float4 chunks are either inside the map or beyond its end.
                // segp <= 132, so a lane owns one chunk column; pointers advance by constants (no per-copy index math)
                if (4 * lane < g.segp) {
                    const bool valid = g.j_begin + 4 * lane < g.iw;
                    const float* sp = inb + warp * cstride + 4 * lane;
                    float* dp = dst + warp * g.segp + 4 * lane;
                    for (int ci = warp; ci < C; ci += NW, sp += NW * cstride, dp += NW * g.segp) cp_async16z(dp, sp, valid);
                }
            } else {
                for (int px = lane; px < g.segp; px += 32) {
                    const bool valid = g.j_begin + px < g.iw;
                    const float* sp = inb + warp * cstride + px;
                    float* dp = dst + warp * g.segp + px;
                    for (int ci = warp; ci < C; ci += NW, sp += NW * cstride, dp += NW * g.segp) cp_async4(dp, sp, valid);
                }
            }
            // phase matrices W[ty][tx0..tx0+ntx): pre-packed (pcnn_upsample_merge_pack_kernel) as (tx, 8-channel group)
            // blocks of 8*C + 4 floats, so ONE bulk copy (TMA engine, tracked by the buffer's mbarrier) brings the
            // whole stage in the bank-staggered layout the FMA loop reads
            float* wdst = dst + p.tile_floats;
            const int ntx = p.st_ntx[st], nblk = ntx * ncg;
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)(nblk * wblk * 4);
                mbar_expect_tx(s_bar + (step & 1), bytes);
                bulk_copy_g2s(wdst, p.dc_w[br] + (long long)(g.ty * g.s + p.st_tx0[st]) * ncg * wblk, bytes, s_bar + (step & 1));
            }
            // bias (zeros when the branch has none) behind the phase matrices: the gather reads it from shared memory
            if (tid < C) cp_async4(wdst + nblk * wblk + tid, p.dc_b[br] ? p.dc_b[br] + tid : p.dc_w[br], p.dc_b[br] != nullptr);
        } else {
            // all resize branches: the tiny source maps, then this row's / these columns' interpolation tables
            for (int r = 0; r < p.n_rs; ++r) {
                const int taps = p.rs_taps[r];
                const int nsrc = C * p.rs_ih[r] * p.rs_iw[r];
                const float* src = p.rs_in[r] + (long long)b * nsrc;
                float* sdst = dst + p.rs_off[r];
                for (int e = tid; e < (nsrc >> 2); e += NTHR) cp_async16z(sdst + 4 * e, src + 4 * e, true);
                float* tab = dst + p.rs_tab[r];                      // [iy(4) | wy(4) | ix(SEG_W*4) | wx(SEG_W*4)]
                if (tid < taps) {
                    cp_async4(tab + tid, reinterpret_cast<const float*>(p.rs_iy[r] + Y * taps + tid), true);
                    cp_async4(tab + 4 + tid, p.rs_wy[r] + Y * taps + tid, true);
                }
                const int left = (p.W - X0) * taps;                  // valid table entries from X0 on
                for (int e = tid; e < ((SEG_W * taps) >> 2); e += NTHR) {
                    const int bytes = min(max((left - 4 * e) * 4, 0), 16);
                    cp_async16p(tab + 8 + 4 * e, reinterpret_cast<const float*>(p.rs_ix[r] + X0 * taps) + 4 * e, bytes);
                    cp_async16p(tab + 8 + SEG_W * 4 + 4 * e, p.rs_wx[r] + X0 * taps + 4 * e, bytes);
                }
            }
        }
    };

    uint32_t bar_phase[2] = {0u, 0u};
    float sum[32];                         // running sums of this thread's pixel (threads < 256), all channels
    int cb = 0, cY = 0, cX0 = 0, nb = 0, nY = 0, nX0 = 0;     // coordinates of the current and of the next step's unit
    int st = 0, kk = 0;
    if (total > 0) { unit_coords(0, cb, cY, cX0); issue_load(0, 0, cb, cY, cX0); }
#ifdef PCNN_UM_PROFILE
    long long t_prev = clock64();
#endif
#pragma unroll 1
    for (int step = 0; step < total; ++step, st = (st + 1 == nstages) ? 0 : st + 1) {
        if (step > 0 && st == 0) { ++kk; cb = nb; cY = nY; cX0 = nX0; }
        cp_async_commit_wait_all();
        UM_T(0);
        __syncthreads();                   // operands of this step landed; everyone is done with the other buffer and with s_st
        UM_T(1);
        if (p.st_branch[st] >= 0) {       // phase matrices landed (the barrier of a buffer flips once per deconv stage using it)
            mbar_wait(s_bar + (step & 1), bar_phase[step & 1]);
            bar_phase[step & 1] ^= 1u;
        }
        UM_T(8);
        const int st_next = (st + 1 == nstages) ? 0 : st + 1;
        if (step + 1 < total) {
            if (st_next == 0) unit_coords(kk + 1, nb, nY, nX0); else { nb = cb; nY = cY; nX0 = cX0; }
            issue_load(step + 1, st_next, nb, nY, nX0);
        }
        UM_T(2);
        const int b = cb, Y = cY, X0 = cX0;
        const int X = X0 + tid;
        const bool owner = tid < SEG_W && X < p.W;
        if (st == 0) {
#pragma unroll
            for (int c = 0; c < 32; ++c) sum[c] = 0.f;
        }
        const float* s_x = s_buf[step & 1];
        const int br = p.st_branch[st];
        if (br >= 0) {
            // ---------------- transpose-conv branch, column phases [tx0, tx0 + ntx)
            const int d = br;
            const DcGeom g = dc_geom(p, d, X0, Y);
            const int s = g.s, segp = g.segp, pgs = g.pgs, ps = p.dc_ps[d];
            const int tx0 = p.st_tx0[st], ntx = p.st_ntx[st];
            const float* s_w = s_x + p.tile_floats;
            const int items = ntx * 4 * pgs;
            for (int it = tid; it < items; it += NTHR) {
                const int tx = it / (4 * pgs);
                const int rem = it - tx * 4 * pgs;
                const int cg = rem / pgs, pg = rem - cg * pgs;
                if (cg * 8 >= C) continue;          // C is a multiple of 8
                float acc[8][4];
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[c][q] = 0.f;
                const float* wp = s_w + (tx * ncg + cg) * wblk;
                const float* xp = s_x + pg * 4;
#pragma unroll 2
                for (int ci = 0; ci < C; ci += 4) {
                    float4 x4[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) x4[kk] = *reinterpret_cast<const float4*>(xp + (ci + kk) * segp);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 w = *reinterpret_cast<const float4*>(wp + c * C + ci);
                        acc[c][0] = fmaf(w.x, x4[0].x, acc[c][0]); acc[c][1] = fmaf(w.x, x4[0].y, acc[c][1]);
                        acc[c][2] = fmaf(w.x, x4[0].z, acc[c][2]); acc[c][3] = fmaf(w.x, x4[0].w, acc[c][3]);
                        acc[c][0] = fmaf(w.y, x4[1].x, acc[c][0]); acc[c][1] = fmaf(w.y, x4[1].y, acc[c][1]);
                        acc[c][2] = fmaf(w.y, x4[1].z, acc[c][2]); acc[c][3] = fmaf(w.y, x4[1].w, acc[c][3]);
                        acc[c][0] = fmaf(w.z, x4[2].x, acc[c][0]); acc[c][1] = fmaf(w.z, x4[2].y, acc[c][1]);
                        acc[c][2] = fmaf(w.z, x4[2].z, acc[c][2]); acc[c][3] = fmaf(w.z, x4[2].w, acc[c][3]);
                        acc[c][0] = fmaf(w.w, x4[3].x, acc[c][0]); acc[c][1] = fmaf(w.w, x4[3].y, acc[c][1]);
                        acc[c][2] = fmaf(w.w, x4[3].z, acc[c][2]); acc[c][3] = fmaf(w.w, x4[3].w, acc[c][3]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(s_st + tx * ps + (cg * 8 + c) * segp + pg * 4) =
                        make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
            }
            UM_T(3);
            __syncthreads();
            UM_T(4);
            if (owner) {
                const int jx = X + p.dc_pbw[d];
                const int j = fast_div(jx, p.dc_magic[d]), tx = jx - j * s - tx0;
                const float* gsrc = s_st + tx * ps + (j - g.j_begin);
                const float* bias = s_w + ntx * ncg * wblk;
                const int act = p.dc_act[d];
                if (tx < 0 || tx >= ntx) {
                    // this pixel's column phase belongs to another stage of the branch
                } else if (act == PCNN_ACT_LEAKY_RELU) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (c < C) {
                            const float v = gsrc[c * segp] + bias[c];
                            sum[c] += fmaxf(v, 0.2f * v);
                        }
                    }
                } else if (act == PCNN_ACT_TANH) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (c < C) sum[c] += tanhf(gsrc[c * segp] + bias[c]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (c < C) sum[c] += gsrc[c * segp] + bias[c];
                    }
                }
            }
        } else {
            // ---------------- all resize branches (tiny sources in shared memory): the row interpolation is the same for
            // the whole CTA, so it is done once per branch ([C][iw] values into s_st); each pixel then interpolates along x
            for (int r = 0; r < p.n_rs; ++r) {
                const int ih = p.rs_ih[r], iw = p.rs_iw[r], taps = p.rs_taps[r];
                const float* tab = s_x + p.rs_tab[r];
                int iyv[4];
                float wyv[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {       // unused taps: weight 0 on row 0
                    const bool on = a < taps;
                    iyv[a] = on ? __float_as_int(tab[a]) * iw : 0;
                    wyv[a] = on ? tab[4 + a] : 0.f;
                }
                float* rowbuf = s_st + p.rs_st[r];
                for (int e = tid; e < C * iw; e += NTHR) {
                    const int c = e / iw, xs = e - c * iw;
                    const float* src = s_x + p.rs_off[r] + c * ih * iw + xs;
                    float acc = src[iyv[0]] * wyv[0];
                    acc = fmaf(src[iyv[1]], wyv[1], acc);
                    if (taps > 2) { acc = fmaf(src[iyv[2]], wyv[2], acc); acc = fmaf(src[iyv[3]], wyv[3], acc); }
                    rowbuf[e] = acc;
                }
            }
            __syncthreads();
            if (owner) {
                for (int r = 0; r < p.n_rs; ++r) {
                    const int iw = p.rs_iw[r], taps = p.rs_taps[r];
                    const float* tab = s_x + p.rs_tab[r];
                    int ix[4];
                    float wx[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const bool on = a < taps;
                        ix[a] = on ? __float_as_int(tab[8 + tid * taps + a]) : 0;
                        wx[a] = on ? tab[8 + SEG_W * 4 + tid * taps + a] : 0.f;     // unused taps: weight 0 on element 0
                    }
                    const float* rowbuf = s_st + p.rs_st[r];
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (c < C) {
                            const float* row = rowbuf + c * iw;
                            float acc = row[ix[0]] * wx[0];
                            acc = fmaf(row[ix[1]], wx[1], acc);
                            if (taps > 2) { acc = fmaf(row[ix[2]], wx[2], acc); acc = fmaf(row[ix[3]], wx[3], acc); }
                            sum[c] += acc;
                        }
                    }
                }
            }
        }
        UM_T(br >= 0 ? 5 : 7);
        if (st != nstages - 1 || !owner) continue;

    // ---------------- write-out: one pixel per thread, 16-byte units of the BLK8 layout
    const int Hp = p.H + 2 * HALO, P = p.W + 2 * HALO;
    const size_t plane_px = (size_t)Hp * P;
    const size_t pix = (size_t)(Y + HALO) * P + (X + HALO);
    const int planes = (C + 7) / 8;
    float lo[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) sum[c] = (c < C) ? sum[c] * p.alpha : 0.f;
#pragma unroll
    for (int pl = 0; pl < 4; ++pl) {
        if (pl < planes) {
            const uint4 hv = make_uint4(pack_h2(sum[8 * pl], sum[8 * pl + 1]), pack_h2(sum[8 * pl + 2], sum[8 * pl + 3]),
                                        pack_h2(sum[8 * pl + 4], sum[8 * pl + 5]), pack_h2(sum[8 * pl + 6], sum[8 * pl + 7]));
            *reinterpret_cast<uint4*>(p.out + (((size_t)b * p.c8_total + p.plane0 + pl) * plane_px + pix) * 8) = hv;
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hw[k]));
                lo[8 * pl + 2 * k] = sum[8 * pl + 2 * k] - back.x;
                lo[8 * pl + 2 * k + 1] = sum[8 * pl + 2 * k + 1] - back.y;
            }
            if (p.mode == 2) {
                const uint4 lv = make_uint4(pack_h2(lo[8 * pl], lo[8 * pl + 1]), pack_h2(lo[8 * pl + 2], lo[8 * pl + 3]),
                                            pack_h2(lo[8 * pl + 4], lo[8 * pl + 5]), pack_h2(lo[8 * pl + 6], lo[8 * pl + 7]));
                *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out_lo) + (((size_t)b * p.c8_total + p.plane0 + pl) * plane_px + pix) * 8) = lv;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) lo[8 * pl + k] = 0.f;
        }
    }
    if (p.mode == 3) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            if (2 * g < planes) {
                const float* v = sum + 16 * g;
                const float* l = lo + 16 * g;
                const uint4 qv = make_uint4(e4m3x4(v[0], v[1], v[2], v[3]), e4m3x4(v[4], v[5], v[6], v[7]),
                                            e4m3x4(v[8], v[9], v[10], v[11]), e4m3x4(v[12], v[13], v[14], v[15]));
                const uint4 lv = make_uint4(e4m3x4(l[0] * LO_SCALE, l[1] * LO_SCALE, l[2] * LO_SCALE, l[3] * LO_SCALE),
                                            e4m3x4(l[4] * LO_SCALE, l[5] * LO_SCALE, l[6] * LO_SCALE, l[7] * LO_SCALE),
                                            e4m3x4(l[8] * LO_SCALE, l[9] * LO_SCALE, l[10] * LO_SCALE, l[11] * LO_SCALE),
                                            e4m3x4(l[12] * LO_SCALE, l[13] * LO_SCALE, l[14] * LO_SCALE, l[15] * LO_SCALE));
                uint8_t* q = p.out_lo + (((size_t)b * p.c8_total + p.plane0 + 2 * g) * plane_px + pix) * 16;
                if (p.plane0 + 2 * g == p.tail_pl) {      // lone plane: [x x 8 | remainder x 8] in its own q plane
                    *reinterpret_cast<uint4*>(q) = make_uint4(qv.x, qv.y, lv.x, lv.y);
                } else {
                    *reinterpret_cast<uint4*>(q) = qv;
                    *reinterpret_cast<uint4*>(q + plane_px * 16) = lv;
                }
            }
        }
    }
    UM_T(6);
    }   // step loop
}

// Keras deconv kernel [s][s][Cout=C][Cin=C] -> [s][s][C/8][8*C + 4]: (phase, 8-output-channel group) blocks of 8 rows x C
// input channels, each followed by 4 pad floats so consecutive blocks start 16 B apart in the shared-memory banks
__global__ void pack_kernel(const float* __restrict__ k, float* __restrict__ out, int C, long long nblocks) {
    const int wblk = 8 * C + 4;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < nblocks * wblk; idx += (long long)gridDim.x * blockDim.x) {
        const long long blk = idx / wblk;
        const int e = (int)(idx - blk * wblk);
        out[idx] = e < 8 * C ? k[blk * 8 * C + e] : 0.f;
    }
}

}  // namespace um
}  // namespace pcnn

using namespace pcnn;
using namespace pcnn::um;

extern "C" size_t pcnn_upsample_merge_packed_floats(int stride, int C) {
    if (stride < 1 || C < 8 || (C % 8)) return 0;
    return (size_t)stride * stride * (C / 8) * (8 * C + 4);
}

extern "C" int pcnn_upsample_merge_pack_kernel(const float* kernel, float* packed, int stride, int C, void* stream) {
    PCNN_CHECK_ARG(kernel && packed && stride >= 1 && stride <= 32 && C >= 8 && C <= 32 && (C % 8) == 0, "upsample_merge_pack_kernel: bad argument");
    const long long nblocks = (long long)stride * stride * (C / 8);
    pack_kernel<<<(int)std::min<long long>((nblocks * (8 * C + 4) + 255) / 256, 1184), 256, 0, (cudaStream_t)stream>>>(kernel, packed, C, nblocks);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_upsample_merge_blk8(int n_deconv, const float* const* dc_in, const float* const* dc_kernel,
                                        const float* const* dc_bias, const int* dc_stride, const int* dc_ih, const int* dc_iw,
                                        const int* dc_act, int n_resize, const float* const* rs_in,
                                        const int32_t* const* rs_iy, const float* const* rs_wy, const int32_t* const* rs_ix,
                                        const float* const* rs_wx, const int* rs_taps, const int* rs_ih, const int* rs_iw,
                                        float alpha, void* out, void* out_lo, int mode, int B, int C, int H, int W,
                                        int c_total, int c_offset, void* stream) {
    PCNN_CHECK_ARG(n_deconv >= 0 && n_deconv <= MAXB && n_resize >= 0 && n_resize <= MAXB && n_deconv + n_resize > 0,
                   "upsample_merge_blk8: between 1 and %d branches of each kind", MAXB);
    PCNN_CHECK_ARG(out && mode >= 1 && mode <= 3 && (mode == 1 || out_lo), "upsample_merge_blk8: bad destination / precision mode");
    PCNN_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && H <= 65535 && W > 0, "upsample_merge_blk8: bad shape");
    PCNN_CHECK_ARG(C >= 8 && C <= 32 && (C % 8) == 0, "upsample_merge_blk8: channels %d must be a multiple of 8, <= 32", C);
    PCNN_CHECK_ARG((c_offset % 16) == 0 && c_offset + C <= ((c_total + 15) / 16) * 16, "upsample_merge_blk8: channel offset must be a multiple of 16 inside the buffer");
    const int np_total = (c_total + 7) / 8;
    const int tail_pl = (mode == 3 && (np_total & 1)) ? np_total - 1 : -1;
    PCNN_CHECK_ARG(mode != 3 || (C % 16) == 0 || c_offset / 8 + C / 8 - 1 == tail_pl,
                   "upsample_merge_blk8: precision mode 3 writes whole 16-channel groups or ends on the tensor's lone last plane (C %d at %d of %d)", C, c_offset, c_total);
    Params p;
    p.n_dc = n_deconv; p.n_rs = n_resize; p.alpha = alpha;
    p.out = (__half*)out; p.out_lo = (uint8_t*)out_lo; p.mode = mode;
    p.C = C; p.H = H; p.W = W; p.c8_total = ((c_total + 15) / 16) * 2; p.plane0 = c_offset / 8; p.tail_pl = tail_pl;
    p.B = B;
    // stage order: weight blocks (s*C*C floats) alternate large / small so that the two operand buffers are
    // [largest, 3rd, 5th ...] and [smallest, ...]: the even-step buffer holds the big blocks, the odd one the small
    int order[MAXB], nd = n_deconv;
    for (int d = 0; d < nd; ++d) order[d] = d;
    std::sort(order, order + nd, [&](int a, int b2) { return dc_stride[a] > dc_stride[b2]; });
    int seq[MAXB];
    for (int lo = 0, hi = nd - 1, k = 0; lo <= hi; ) { seq[k++] = order[lo++]; if (lo <= hi) seq[k++] = order[hi--]; }
    int tile = 0, st = 0, rs_need = 0, rs_st = 0;
    size_t need[2] = {0, 0};
    for (int k = 0; k < nd; ++k) {
        const int d = seq[k];
        const int s = dc_stride[d], ih = dc_ih[d], iw = dc_iw[d];
        PCNN_CHECK_ARG(dc_in[d] && dc_kernel[d] && s >= 1 && s <= 32, "upsample_merge_blk8: deconv branch %d: bad argument", d);
        PCNN_CHECK_ARG((reinterpret_cast<uintptr_t>(dc_kernel[d]) % 16) == 0, "upsample_merge_blk8: deconv kernel %d must be 16-byte aligned", d);
        PCNN_CHECK_ARG(ceil_div(H, s) == ih && ceil_div(W, s) == iw,
                       "upsample_merge_blk8: output (%d,%d) inconsistent with input (%d,%d) at stride %d (TF raises)", H, W, ih, iw, s);
        p.dc_al16[k] = (reinterpret_cast<uintptr_t>(dc_in[d]) % 16) == 0;
        p.dc_magic[k] = (uint32_t)((0x100000000ull + (unsigned)s - 1) / (unsigned)s);
        p.dc_in[k] = dc_in[d]; p.dc_w[k] = dc_kernel[d]; p.dc_b[k] = dc_bias ? dc_bias[d] : nullptr;
        p.dc_s[k] = s; p.dc_ih[k] = ih; p.dc_iw[k] = iw; p.dc_act[k] = dc_act[d];
        p.dc_pbh[k] = std::max((ih - 1) * s + s - H, 0) / 2;
        p.dc_pbw[k] = std::max((iw - 1) * s + s - W, 0) / 2;
        const int segp_max = (((SEG_W - 1) / s + 2) + 3) & ~3;           // low-res pixels touching one row segment
        int ps = 32 * segp_max + (32 + s - 1) / s;                        // phases land ~32/s banks apart
        ps = (ps + 3) & ~3;
        p.dc_ps[k] = ps;
        tile = std::max(tile, C * segp_max);
        st = std::max(st, std::min(s, TXG) * ps);
    }
    for (int r = 0; r < n_resize; ++r) {
        PCNN_CHECK_ARG(rs_in[r] && rs_iy[r] && rs_wy[r] && rs_ix[r] && rs_wx[r] && rs_taps[r] >= 1 && rs_taps[r] <= 4 && rs_ih[r] > 0 && rs_iw[r] > 0,
                       "upsample_merge_blk8: resize branch %d: bad argument", r);
        PCNN_CHECK_ARG((long long)C * rs_ih[r] * rs_iw[r] <= 8192, "upsample_merge_blk8: resize source %dx%d too large for the fused kernel", rs_ih[r], rs_iw[r]);
        p.rs_in[r] = rs_in[r]; p.rs_iy[r] = rs_iy[r]; p.rs_wy[r] = rs_wy[r]; p.rs_ix[r] = rs_ix[r]; p.rs_wx[r] = rs_wx[r];
        p.rs_taps[r] = rs_taps[r]; p.rs_ih[r] = rs_ih[r]; p.rs_iw[r] = rs_iw[r];
        PCNN_CHECK_ARG(((reinterpret_cast<uintptr_t>(rs_in[r]) | reinterpret_cast<uintptr_t>(rs_ix[r]) | reinterpret_cast<uintptr_t>(rs_wx[r])) % 16) == 0,
                       "upsample_merge_blk8: resize branch %d: source and tables must be 16-byte aligned", r);
        p.rs_off[r] = rs_need;
        p.rs_tab[r] = rs_need + ((C * rs_ih[r] * rs_iw[r] + 3) & ~3);
        rs_need = p.rs_tab[r] + 8 + 2 * SEG_W * 4;
        p.rs_st[r] = rs_st;
        rs_st += (C * rs_iw[r] + 3) & ~3;
    }
    st = std::max(st, rs_st);
    tile = (tile + 3) & ~3;
    st = (st + 3) & ~3;
    // stage table: every deconv branch contributes ceil(s / TXG) stages, then the resize branches
    int nstages = 0;
    size_t wsize[MAXS];
    for (int k = 0; k < nd; ++k)
        for (int tx0 = 0; tx0 < p.dc_s[k]; tx0 += TXG) {
            PCNN_CHECK_ARG(nstages < MAXS, "upsample_merge_blk8: too many stages");
            p.st_branch[nstages] = (signed char)k; p.st_tx0[nstages] = (unsigned char)tx0;
            p.st_ntx[nstages] = (unsigned char)std::min(TXG, p.dc_s[k] - tx0);
            wsize[nstages] = (size_t)tile + (size_t)p.st_ntx[nstages] * (C / 8) * (8 * C + 4) + 32;
            ++nstages;
        }
    if (n_resize > 0) {
        PCNN_CHECK_ARG(nstages < MAXS, "upsample_merge_blk8: too many stages");
        p.st_branch[nstages] = (signed char)-1; p.st_tx0[nstages] = 0; p.st_ntx[nstages] = 0;
        wsize[nstages] = (size_t)rs_need;
        ++nstages;
    }
    p.n_stages = nstages;
    for (int k = 0; k < nstages; ++k) {
        const int par = (nstages % 2 == 0) ? (k & 1) : 0;       // odd stage count: step parity drifts, both buffers get the maximum
        need[par] = std::max(need[par], wsize[k]);
    }
    if (nstages % 2) need[1] = need[0];
    p.tile_floats = tile; p.st_floats = st; p.buf_floats[0] = (int)need[0]; p.buf_floats[1] = (int)need[1];
    need[0] = (need[0] + 3) & ~(size_t)3; need[1] = (need[1] + 3) & ~(size_t)3;
    p.buf_floats[0] = (int)need[0]; p.buf_floats[1] = (int)need[1];
    const size_t smem = ((size_t)4 + st + need[0] + need[1]) * sizeof(float);
    PCNN_CHECK_ARG(smem <= 227 * 1024, "upsample_merge_blk8: shared-memory plan too large (%zu bytes)", smem);
    PCNN_CHECK_CUDA(cudaFuncSetAttribute(upsample_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    PCNN_CHECK_CUDA(cudaGetDevice(&dev));
    PCNN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long units = (long long)B * H * ceil_div(W, SEG_W);
    PCNN_CHECK_ARG(units < (1ll << 30), "upsample_merge_blk8: too many row segments");
    const int ctas_per_sm = std::max(1, std::min(4, (int)((227 * 1024) / (smem + 1024))));
    const int grid = (int)std::min<long long>(units, (long long)sms * ctas_per_sm);
    upsample_merge_kernel<<<grid, NTHR, smem, (cudaStream_t)stream>>>(p);
    PCNN_CHECK_LAUNCH();
#ifdef PCNN_UM_PROFILE
    {
        unsigned long long h[16];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_um_prof, sizeof(h));
        const char* names[9] = {"cp.async wait", "barrier (top)", "issue next loads", "compute", "barrier (mid)", "deconv gather", "write-out", "resize stage", "weights mbarrier"};
        unsigned long long tot = 0; for (int i = 0; i < 9; ++i) tot += h[i];
        for (int i = 0; i < 9; ++i) fprintf(stderr, "um phase %-18s %12llu cycles %5.1f%%\n", names[i], h[i], 100.0 * h[i] / (tot ? tot : 1));
        unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_um_prof, z, sizeof(z));
    }
#endif
    return PCNN_OK;
}
