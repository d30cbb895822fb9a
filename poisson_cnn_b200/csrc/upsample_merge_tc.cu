// Tensor-core version of the fused upsample + merge of the HPNN bottleneck branches (see upsample_merge.cu for the
// operator: models/Homogeneous_Poisson_NN_Legacy.py:226-233, blocks/bottleneck_block.py:57-118,
// layers/deconvupscale.py:100-109, layers/Upsample.py:57 of the reference).
//
// The k == stride transpose convolutions are, per output row Y and column phase tx, plain 32x32 matrix products
//     P[j, co] = sum_ci in[i(Y), j, ci] * W[ty(Y), tx, co, ci]        (output pixel x = j*s + tx - pad_left)
// so they run on the tensor cores (mma.sync m16n8k16, fp16 operands, fp32 accumulation) straight from the BLK8 fp16
// tensors the branch convolutions produce: no fp32 copy of the branch outputs exists any more.  The FP32-FMA kernel in
// upsample_merge.cu spends 5 x 1024 FMAs per output pixel on this (3.9 ms per 64 samples at 17 % of the FMA pipe).
//
// Work split: persistent CTAs walk chunks = (output row Y, 256-pixel segment, group of CH samples); the s phase
// matrices of row phase ty of every branch (76 KB for strides 2,3,4,8,16) are staged ONCE per chunk, the low-res input
// rows (22 KB) and the tiny resize sources per sample, double-buffered, all with cp.async.bulk + mbarriers from one
// producer warp.  Each of the 8 compute warps OWNS 32 output pixels: it sums the resize branches (lane = pixel), then
// for every branch runs the M-tiles of low-res pixels that touch its range and adds bias + activation into its private
// [32 px][32 ch] fp32 tile in shared memory (the accumulator fragments of branches with different strides map to
// different pixels, so the cross-branch sum cannot stay in registers), and finally writes its 32 pixels once in the
// BLK8 operand layout of the consumer (fp16 + remainder / e4m3 planes).  No block-wide barrier in steady state.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <algorithm>

#include "pcnn_common.cuh"

namespace pcnn {
namespace umt {

constexpr int HALO = 7;
constexpr int MAXB = 8;          // branches of each kind
constexpr int SEG_W = 256;       // output pixels per segment = 8 compute warps x 32
constexpr int NWARP = 8;
constexpr int NTHR = (NWARP + 1) * 32;
constexpr int CH = 8;            // samples per chunk (the staged phase matrices are reused for all of them)
constexpr int C = 32;            // channels (the shipped architecture; other widths use the FP32 kernel)
constexpr int WROW = 36;         // halves per weight row: 32 input channels + 4 pad -> conflict-free B-fragment loads
constexpr int WMAT = C * WROW;   // halves per phase matrix
constexpr float LO_SCALE = 2048.f;
constexpr unsigned long long SPIN_LIMIT_NS = 4000000000ull;

struct Params {
    int n_dc, n_rs;
    const __half* dc_in[MAXB];   // BLK8 fp16 [B][4][ih+14][iw+14][8]
    const __half* dc_w[MAXB];    // fp16 [s][s][32][36]  (pcnn_upsample_merge_tc_pack_kernel)
    const float* dc_b[MAXB];     // [32] or null
    int dc_s[MAXB], dc_ih[MAXB], dc_iw[MAXB], dc_pbh[MAXB], dc_pbw[MAXB], dc_act[MAXB];
    int dc_segp[MAXB];           // staged low-res pixels per segment (fixed: covers every segment)
    int dc_aoff[MAXB];           // byte offset of the branch's row tile inside an A buffer
    int dc_woff[MAXB];           // byte offset of the branch's phase matrices inside the W buffer
    uint32_t dc_magic[MAXB];     // ceil(2^32 / stride)
    const float* rs_in[MAXB];    // [B][32][ih][iw] fp32
    const int* rs_iy[MAXB]; const float* rs_wy[MAXB]; const int* rs_ix[MAXB]; const float* rs_wx[MAXB];
    int rs_taps[MAXB], rs_ih[MAXB], rs_iw[MAXB];
    int rs_soff[MAXB];           // byte offset of the source map inside an R buffer
    int rs_roff[MAXB];           // float offset of the interpolated row inside a warp's row buffer
    int a_bytes, w_bytes, r_bytes, rowbuf_floats;   // sizes of one A buffer, the W buffer, one R buffer, one warp's row buffer
    float alpha;
    __half* out; uint8_t* out_lo;
    int mode;                    // precision mode of the destination (1, 2, 3: see pcnn_conv2d_tc)
    int B, H, W, c8_total, plane0, tail_pl;
};

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
            printf("pcnn upsample_merge_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int fast_div(int n, uint32_t magic) { return (int)__umulhi((uint32_t)n, magic); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {
    const uint32_t lo = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}
// channel rotation of pixel px inside a warp tile (units of 8 channels): spreads the rows an accumulator fragment touches
// (px = s*j + const for every stride s) and the rows of consecutive pixels over the four 8-word bank groups
__device__ __forceinline__ int rot8(int px) { return ((px + (px >> 2) + (px >> 4)) & 3) * 8; }

__global__ void __launch_bounds__(NTHR, 1) upsample_merge_tc_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    // carve-up: [W][A0][A1][R0][R1][warp tiles 8 x 4 KB][warp row buffers][mbarriers]
    uint8_t* s_w = smem;
    uint8_t* s_a[2] = {s_w + p.w_bytes, s_w + p.w_bytes + p.a_bytes};
    uint8_t* s_r[2] = {s_a[1] + p.a_bytes, s_a[1] + p.a_bytes + p.r_bytes};
    float* s_tile = reinterpret_cast<float*>(s_r[1] + p.r_bytes);
    float* s_row = s_tile + NWARP * 32 * C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_row + NWARP * p.rowbuf_floats);
    uint64_t* in_full = bars;          // [2] operands of a sample landed
    uint64_t* in_empty = bars + 2;     // [2] all compute warps are done with them
    uint64_t* w_full = bars + 4;       // phase matrices of the chunk landed
    uint64_t* w_empty = bars + 5;      // all compute warps are done with the chunk

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int segs = (p.W + SEG_W - 1) / SEG_W;
    const int n_bc = (p.B + CH - 1) / CH;
    const int chunks = p.H * segs * n_bc;
    if (threadIdx.x == 0) {
        mbar_init(in_full, 1); mbar_init(in_full + 1, 1);
        mbar_init(in_empty, NWARP); mbar_init(in_empty + 1, NWARP);
        mbar_init(w_full, 1); mbar_init(w_empty, NWARP);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto chunk_coords = [&](int k, int& Y, int& X0, int& b0, int& nb) {
        const int bc = k % n_bc;
        const int t = k / n_bc;
        X0 = (t % segs) * SEG_W; Y = t / segs;
        b0 = bc * CH; nb = min(CH, p.B - b0);
    };

    if (warp == NWARP) {
        // ================= producer: bulk copies of the chunk's phase matrices and of every sample's operands
        if (lane == 0) {
            uint32_t unit = 0, cidx = 0;
            for (int k = blockIdx.x; k < chunks; k += gridDim.x, ++cidx) {
                int Y, X0, b0, nb;
                chunk_coords(k, Y, X0, b0, nb);
                mbar_wait(w_empty, (cidx & 1) ^ 1);
                mbar_expect_tx(w_full, (uint32_t)p.w_bytes);
                for (int d = 0; d < p.n_dc; ++d) {
                    const int s = p.dc_s[d];
                    const int jy = Y + p.dc_pbh[d];
                    const int i = fast_div(jy, p.dc_magic[d]), ty = jy - i * s;
                    bulk_copy_g2s(s_w + p.dc_woff[d], p.dc_w[d] + (size_t)ty * s * WMAT, (uint32_t)(s * WMAT * 2), w_full);
                }
                for (int q = 0; q < nb; ++q, ++unit) {
                    const int buf = unit & 1, b = b0 + q;
                    mbar_wait(in_empty + buf, ((unit >> 1) & 1) ^ 1);
                    mbar_expect_tx(in_full + buf, (uint32_t)(p.a_bytes + p.r_bytes));
                    for (int d = 0; d < p.n_dc; ++d) {
                        const int Hp = p.dc_ih[d] + 2 * HALO, P = p.dc_iw[d] + 2 * HALO;
                        const int i = fast_div(Y + p.dc_pbh[d], p.dc_magic[d]);
                        const int jb = fast_div(X0 + p.dc_pbw[d], p.dc_magic[d]);
                        const uint32_t bytes = (uint32_t)p.dc_segp[d] * 16u;
                        const __half* src = p.dc_in[d] + (((size_t)b * 4 * Hp + (i + HALO)) * P + (jb + HALO)) * 8;
                        for (int pl = 0; pl < 4; ++pl)
                            bulk_copy_g2s(s_a[buf] + p.dc_aoff[d] + pl * bytes, src + (size_t)pl * Hp * P * 8, bytes, in_full + buf);
                    }
                    for (int r = 0; r < p.n_rs; ++r) {
                        const uint32_t bytes = (uint32_t)(C * p.rs_ih[r] * p.rs_iw[r] * 4);
                        bulk_copy_g2s(s_r[buf] + p.rs_soff[r], p.rs_in[r] + (size_t)b * C * p.rs_ih[r] * p.rs_iw[r], bytes, in_full + buf);
                    }
                }
            }
        }
        return;
    }

    // ================= compute warps: warp w owns output pixels [X0 + 32 w, X0 + 32 w + 32)
    float* tile = s_tile + warp * 32 * C;                  // [32 px][32 ch] fp32, channel octets rotated by rot8(px)
    float* rowbuf = s_row + warp * p.rowbuf_floats;
    const int g = lane >> 2, t = lane & 3;
    const int Hp = p.H + 2 * HALO, P = p.W + 2 * HALO;
    const size_t plane_px = (size_t)Hp * P;
    uint32_t unit = 0, cidx = 0;
    for (int k = blockIdx.x; k < chunks; k += gridDim.x, ++cidx) {
        int Y, X0, b0, nb;
        chunk_coords(k, Y, X0, b0, nb);
        const int xw = X0 + 32 * warp;                     // first pixel of this warp
        const int X = xw + lane;                           // this lane's pixel (resize sum, write-out)
        const bool live_warp = xw < p.W;
        // per-chunk resize tables of this lane's pixel (same for every sample of the chunk)
        mbar_wait(w_full, cidx & 1);
        for (int q = 0; q < nb; ++q, ++unit) {
            const int buf = unit & 1, b = b0 + q;
            mbar_wait(in_full + buf, (unit >> 1) & 1);
            if (live_warp) {
                // ---------------- resize branches: row interpolation (cooperative), then each lane interpolates its pixel
                float sum[C];
#pragma unroll
                for (int c = 0; c < C; ++c) sum[c] = 0.f;
                for (int r = 0; r < p.n_rs; ++r) {
                    const int ih = p.rs_ih[r], iw = p.rs_iw[r], taps = p.rs_taps[r];
                    const float* src = reinterpret_cast<const float*>(s_r[buf] + p.rs_soff[r]);
                    int iyv[4]; float wyv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const bool on = a < taps;
                        iyv[a] = on ? __ldg(p.rs_iy[r] + Y * taps + a) * iw : 0;
                        wyv[a] = on ? __ldg(p.rs_wy[r] + Y * taps + a) : 0.f;
                    }
                    float* rb = rowbuf + p.rs_roff[r];
                    for (int e = lane; e < C * iw; e += 32) {
                        const int c = e / iw, xs = e - c * iw;
                        const float* sp = src + c * ih * iw + xs;
                        float acc = sp[iyv[0]] * wyv[0];
                        acc = fmaf(sp[iyv[1]], wyv[1], acc);
                        acc = fmaf(sp[iyv[2]], wyv[2], acc);
                        acc = fmaf(sp[iyv[3]], wyv[3], acc);
                        rb[e] = acc;
                    }
                }
                __syncwarp();
                if (X < p.W) {
                    for (int r = 0; r < p.n_rs; ++r) {
                        const int iw = p.rs_iw[r], taps = p.rs_taps[r];
                        int ix[4]; float wx[4];
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const bool on = a < taps;
                            ix[a] = on ? __ldg(p.rs_ix[r] + X * taps + a) : 0;
                            wx[a] = on ? __ldg(p.rs_wx[r] + X * taps + a) : 0.f;
                        }
                        const float* rb = rowbuf + p.rs_roff[r];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const float* row = rb + c * iw;
                            float acc = row[ix[0]] * wx[0];
                            acc = fmaf(row[ix[1]], wx[1], acc);
                            acc = fmaf(row[ix[2]], wx[2], acc);
                            acc = fmaf(row[ix[3]], wx[3], acc);
                            sum[c] += acc;
                        }
                    }
                }
                {   // start value of the warp tile: the resize sum (zeros without resize branches)
                    float* row = tile + lane * C;
                    const int rot = rot8(lane);
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        *reinterpret_cast<float4*>(row + ((8 * o + rot) & 31)) = make_float4(sum[8 * o], sum[8 * o + 1], sum[8 * o + 2], sum[8 * o + 3]);
                        *reinterpret_cast<float4*>(row + ((8 * o + rot) & 31) + 4) = make_float4(sum[8 * o + 4], sum[8 * o + 5], sum[8 * o + 6], sum[8 * o + 7]);
                    }
                }
                __syncwarp();
                // ---------------- transpose-conv branches on the tensor cores
                for (int d = 0; d < p.n_dc; ++d) {
                    const int s = p.dc_s[d], pbw = p.dc_pbw[d], segp = p.dc_segp[d], act = p.dc_act[d];
                    const uint32_t mg = p.dc_magic[d];
                    const int jb = fast_div(X0 + pbw, mg);                               // first staged low-res pixel
                    const int j_lo = fast_div(xw + pbw, mg);
                    const int j_hi = fast_div(min(xw + 31, p.W - 1) + pbw, mg);
                    const uint8_t* sa = s_a[buf] + p.dc_aoff[d];
                    const uint8_t* sw = s_w + p.dc_woff[d];
                    const uint32_t plane_b = (uint32_t)segp * 16u;
                    float bias[4][2];
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        bias[n][0] = p.dc_b[d] ? __ldg(p.dc_b[d] + 8 * n + 2 * t) : 0.f;
                        bias[n][1] = p.dc_b[d] ? __ldg(p.dc_b[d] + 8 * n + 2 * t + 1) : 0.f;
                    }
                    for (int j0 = j_lo; j0 <= j_hi; j0 += 16) {
                        // A fragments of the 16 low-res pixels j0 .. j0+15 (rows beyond the staged tile are clamped: their
                        // results fall outside the warp's pixel range or beyond W and are dropped)
                        const int r0 = min(j0 + g - jb, segp - 1), r1 = min(j0 + g + 8 - jb, segp - 1);
                        uint32_t a[2][4];
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint8_t* base = sa + (2 * ks) * plane_b + 4 * t;
                            a[ks][0] = *reinterpret_cast<const uint32_t*>(base + r0 * 16);
                            a[ks][1] = *reinterpret_cast<const uint32_t*>(base + r1 * 16);
                            a[ks][2] = *reinterpret_cast<const uint32_t*>(base + plane_b + r0 * 16);
                            a[ks][3] = *reinterpret_cast<const uint32_t*>(base + plane_b + r1 * 16);
                        }
                        for (int tx = 0; tx < s; ++tx) {
                            // output pixels of the two accumulator rows of this lane
                            const int x0p = (j0 + g) * s + tx - pbw - xw, x1p = x0p + 8 * s;
                            const uint8_t* wb = sw + (size_t)tx * (WMAT * 2) + g * (WROW * 2) + 4 * t;
                            float acc[4][4];
#pragma unroll
                            for (int n = 0; n < 4; ++n) {
                                acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
                                for (int ks = 0; ks < 2; ++ks) {
                                    const uint8_t* wp = wb + n * 8 * (WROW * 2) + ks * 32;
                                    mma_f16(acc[n], a[ks], *reinterpret_cast<const uint32_t*>(wp), *reinterpret_cast<const uint32_t*>(wp + 16));
                                }
                            }
                            const bool ok0 = x0p >= 0 && x0p < 32 && xw + x0p < p.W;
                            const bool ok1 = x1p >= 0 && x1p < 32 && xw + x1p < p.W;
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int xp = h ? x1p : x0p;
                                if (h ? ok1 : ok0) {
                                    float* row = tile + xp * C;
                                    const int rot = rot8(xp);
#pragma unroll
                                    for (int n = 0; n < 4; ++n) {
                                        float v0 = acc[n][2 * h] + bias[n][0], v1 = acc[n][2 * h + 1] + bias[n][1];
                                        if (act == PCNN_ACT_LEAKY_RELU) { v0 = fmaxf(v0, 0.2f * v0); v1 = fmaxf(v1, 0.2f * v1); }
                                        else if (act == PCNN_ACT_TANH) { v0 = tanhf(v0); v1 = tanhf(v1); }
                                        float2* dst = reinterpret_cast<float2*>(row + ((8 * n + rot) & 31) + 2 * t);
                                        float2 cur = *dst;
                                        cur.x += v0; cur.y += v1;
                                        *dst = cur;
                                    }
                                }
                            }
                            __syncwarp();      // two phases of one branch never touch the same pixel, but the next M-tile / branch may
                        }
                    }
                }
                __syncwarp();
            }
            // operands of this sample are consumed (the tile and registers hold everything from here on)
            if (lane == 0) mbar_arrive(in_empty + buf);
            if (live_warp && X < p.W) {
                // ---------------- write-out: one pixel per lane, 16-byte units of the BLK8 layout
                float sum[C], lo[C];
                {
                    const float* row = tile + lane * C;
                    const int rot = rot8(lane);
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const float4 u = *reinterpret_cast<const float4*>(row + ((8 * o + rot) & 31));
                        const float4 v = *reinterpret_cast<const float4*>(row + ((8 * o + rot) & 31) + 4);
                        sum[8 * o] = u.x * p.alpha; sum[8 * o + 1] = u.y * p.alpha; sum[8 * o + 2] = u.z * p.alpha; sum[8 * o + 3] = u.w * p.alpha;
                        sum[8 * o + 4] = v.x * p.alpha; sum[8 * o + 5] = v.y * p.alpha; sum[8 * o + 6] = v.z * p.alpha; sum[8 * o + 7] = v.w * p.alpha;
                    }
                }
                const size_t pix = (size_t)(Y + HALO) * P + (X + HALO);
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    uint32_t hw[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) hw[e] = pack_h2(sum[8 * pl + 2 * e], sum[8 * pl + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(p.out + (((size_t)b * p.c8_total + p.plane0 + pl) * plane_px + pix) * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                        lo[8 * pl + 2 * e] = sum[8 * pl + 2 * e] - back.x;
                        lo[8 * pl + 2 * e + 1] = sum[8 * pl + 2 * e + 1] - back.y;
                    }
                    if (p.mode == 2) {
                        const uint4 lv = make_uint4(pack_h2(lo[8 * pl], lo[8 * pl + 1]), pack_h2(lo[8 * pl + 2], lo[8 * pl + 3]),
                                                    pack_h2(lo[8 * pl + 4], lo[8 * pl + 5]), pack_h2(lo[8 * pl + 6], lo[8 * pl + 7]));
                        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out_lo) + (((size_t)b * p.c8_total + p.plane0 + pl) * plane_px + pix) * 8) = lv;
                    }
                }
                if (p.mode == 3) {
#pragma unroll
                    for (int gq = 0; gq < 2; ++gq) {
                        const float* v = sum + 16 * gq;
                        const float* l = lo + 16 * gq;
                        const uint4 qv = make_uint4(e4m3x4(v[0], v[1], v[2], v[3]), e4m3x4(v[4], v[5], v[6], v[7]),
                                                    e4m3x4(v[8], v[9], v[10], v[11]), e4m3x4(v[12], v[13], v[14], v[15]));
                        const uint4 lv = make_uint4(e4m3x4(l[0] * LO_SCALE, l[1] * LO_SCALE, l[2] * LO_SCALE, l[3] * LO_SCALE),
                                                    e4m3x4(l[4] * LO_SCALE, l[5] * LO_SCALE, l[6] * LO_SCALE, l[7] * LO_SCALE),
                                                    e4m3x4(l[8] * LO_SCALE, l[9] * LO_SCALE, l[10] * LO_SCALE, l[11] * LO_SCALE),
                                                    e4m3x4(l[12] * LO_SCALE, l[13] * LO_SCALE, l[14] * LO_SCALE, l[15] * LO_SCALE));
                        uint8_t* q8 = p.out_lo + (((size_t)b * p.c8_total + p.plane0 + 2 * gq) * plane_px + pix) * 16;
                        *reinterpret_cast<uint4*>(q8) = qv;
                        *reinterpret_cast<uint4*>(q8 + plane_px * 16) = lv;
                    }
                }
            }
            __syncwarp();      // the tile is rewritten by the next sample's resize stage
        }
        // the chunk's phase matrices are consumed
        if (lane == 0) mbar_arrive(w_empty);
    }
}

// Keras deconv kernel [s][s][Cout=32][Cin=32] fp32 -> fp16 [s][s][32][36] (rows padded to 36 halves)
__global__ void pack_tc_kernel(const float* __restrict__ k, __half* __restrict__ out, long long nrows) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < nrows * WROW; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / WROW;
        const int e = (int)(idx - row * WROW);
        out[idx] = __float2half_rn(e < C ? k[row * C + e] : 0.f);
    }
}

}  // namespace umt
}  // namespace pcnn

using namespace pcnn;
using namespace pcnn::umt;

extern "C" size_t pcnn_upsample_merge_tc_packed_bytes(int stride) {
    return (stride >= 1 && stride <= 32) ? (size_t)stride * stride * WMAT * sizeof(__half) : 0;
}

extern "C" int pcnn_upsample_merge_tc_pack_kernel(const float* kernel, void* packed, int stride, void* stream) {
    PCNN_CHECK_ARG(kernel && packed && stride >= 1 && stride <= 32, "upsample_merge_tc_pack_kernel: bad argument");
    const long long nrows = (long long)stride * stride * C;
    const int grid = (int)std::min<long long>((nrows * WROW + 255) / 256, 148 * 8);
    pack_tc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kernel, (__half*)packed, nrows);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

// Shared memory the kernel needs for these branches (0 = unsupported arguments); callers fall back to
// pcnn_upsample_merge_blk8 when it exceeds 227 KB (large resize sources: grids beyond ~400 pixels a side).
extern "C" size_t pcnn_upsample_merge_tc_smem_bytes(int n_deconv, const int* dc_stride, int n_resize, const int* rs_ih, const int* rs_iw) {
    if (n_deconv < 1 || n_deconv > MAXB || n_resize < 0 || n_resize > MAXB || !dc_stride || (n_resize && (!rs_ih || !rs_iw))) return 0;
    size_t a = 0, w = 0, r = 0, row = 0;
    for (int d = 0; d < n_deconv; ++d) {
        const int s = dc_stride[d];
        if (s < 1 || s > 32) return 0;
        a += (size_t)4 * ((((SEG_W - 1) / s + 2) + 3) & ~3) * 16;
        w += (size_t)s * WMAT * 2;
    }
    for (int q = 0; q < n_resize; ++q) {
        if (rs_ih[q] < 1 || rs_iw[q] < 1) return 0;
        r += (size_t)C * rs_ih[q] * rs_iw[q] * 4;
        row += (size_t)((C * rs_iw[q] + 3) & ~3);
    }
    row = (row + 31) & ~(size_t)31;
    return w + 2 * a + 2 * r + (size_t)NWARP * 32 * C * 4 + (size_t)NWARP * row * 4 + 64;
}

extern "C" int pcnn_upsample_merge_tc_blk8(int n_deconv, const void* const* dc_in, const void* const* dc_wpack,
                                           const float* const* dc_bias, const int* dc_stride, const int* dc_ih, const int* dc_iw,
                                           const int* dc_act, int n_resize, const float* const* rs_in,
                                           const int32_t* const* rs_iy, const float* const* rs_wy, const int32_t* const* rs_ix,
                                           const float* const* rs_wx, const int* rs_taps, const int* rs_ih, const int* rs_iw,
                                           float alpha, void* out, void* out_lo, int mode, int B, int H, int W,
                                           int c_total, int c_offset, void* stream) {
    PCNN_CHECK_ARG(n_deconv >= 1 && n_deconv <= MAXB && n_resize >= 0 && n_resize <= MAXB,
                   "upsample_merge_tc_blk8: between 1 and %d transpose-conv branches, at most %d resize branches", MAXB, MAXB);
    PCNN_CHECK_ARG(out && mode >= 1 && mode <= 3 && (mode == 1 || out_lo), "upsample_merge_tc_blk8: bad destination / precision mode");
    PCNN_CHECK_ARG(B > 0 && H > 0 && W > 0, "upsample_merge_tc_blk8: bad shape");
    PCNN_CHECK_ARG((c_offset % 16) == 0 && c_offset + C <= ((c_total + 15) / 16) * 16, "upsample_merge_tc_blk8: channel offset must be a multiple of 16 inside the buffer");
    PCNN_CHECK_ARG(mode != 3 || (((c_total + 7) / 8) & 1) == 0, "upsample_merge_tc_blk8: precision mode 3 needs a destination with an even number of 8-channel planes");
    Params p;
    p.n_dc = n_deconv; p.n_rs = n_resize; p.alpha = alpha;
    p.out = (__half*)out; p.out_lo = (uint8_t*)out_lo; p.mode = mode;
    p.B = B; p.H = H; p.W = W; p.c8_total = ((c_total + 15) / 16) * 2; p.plane0 = c_offset / 8; p.tail_pl = -1;
    int aoff = 0, woff = 0;
    for (int d = 0; d < n_deconv; ++d) {
        const int s = dc_stride[d], ih = dc_ih[d], iw = dc_iw[d];
        PCNN_CHECK_ARG(dc_in[d] && dc_wpack[d] && s >= 1 && s <= 32, "upsample_merge_tc_blk8: branch %d: bad argument", d);
        PCNN_CHECK_ARG((reinterpret_cast<uintptr_t>(dc_wpack[d]) % 16) == 0 && (reinterpret_cast<uintptr_t>(dc_in[d]) % 16) == 0,
                       "upsample_merge_tc_blk8: branch %d: operands must be 16-byte aligned", d);
        PCNN_CHECK_ARG(ceil_div(H, s) == ih && ceil_div(W, s) == iw,
                       "upsample_merge_tc_blk8: output (%d,%d) inconsistent with input (%d,%d) at stride %d (TF raises)", H, W, ih, iw, s);
        p.dc_in[d] = (const __half*)dc_in[d]; p.dc_w[d] = (const __half*)dc_wpack[d]; p.dc_b[d] = dc_bias ? dc_bias[d] : nullptr;
        p.dc_s[d] = s; p.dc_ih[d] = ih; p.dc_iw[d] = iw; p.dc_act[d] = dc_act[d];
        p.dc_pbh[d] = std::max((ih - 1) * s + s - H, 0) / 2;
        p.dc_pbw[d] = std::max((iw - 1) * s + s - W, 0) / 2;
        p.dc_magic[d] = (uint32_t)((0x100000000ull + (unsigned)s - 1) / (unsigned)s);
        p.dc_segp[d] = (((SEG_W - 1) / s + 2) + 3) & ~3;        // low-res pixels touching one segment, rounded up
        p.dc_aoff[d] = aoff; aoff += 4 * p.dc_segp[d] * 16;
        p.dc_woff[d] = woff; woff += s * WMAT * 2;
    }
    int roff = 0, rowf = 0;
    for (int r = 0; r < n_resize; ++r) {
        PCNN_CHECK_ARG(rs_in[r] && rs_iy[r] && rs_wy[r] && rs_ix[r] && rs_wx[r] && rs_taps[r] >= 1 && rs_taps[r] <= 4 && rs_ih[r] > 0 && rs_iw[r] > 0,
                       "upsample_merge_tc_blk8: resize branch %d: bad argument", r);
        PCNN_CHECK_ARG((long long)C * rs_ih[r] * rs_iw[r] <= 8192 && (reinterpret_cast<uintptr_t>(rs_in[r]) % 16) == 0,
                       "upsample_merge_tc_blk8: resize source %dx%d too large for the fused kernel or misaligned", rs_ih[r], rs_iw[r]);
        p.rs_in[r] = rs_in[r]; p.rs_iy[r] = rs_iy[r]; p.rs_wy[r] = rs_wy[r]; p.rs_ix[r] = rs_ix[r]; p.rs_wx[r] = rs_wx[r];
        p.rs_taps[r] = rs_taps[r]; p.rs_ih[r] = rs_ih[r]; p.rs_iw[r] = rs_iw[r];
        p.rs_soff[r] = roff; roff += (C * rs_ih[r] * rs_iw[r] * 4 + 127) & ~127;
        p.rs_roff[r] = rowf; rowf += (C * rs_iw[r] + 3) & ~3;
    }
    p.a_bytes = (aoff + 127) & ~127; p.w_bytes = (woff + 127) & ~127; p.r_bytes = roff; p.rowbuf_floats = (rowf + 31) & ~31;
    // the mbarrier transaction counts are the bytes actually copied
    PCNN_CHECK_ARG(p.a_bytes == aoff && p.w_bytes == woff, "upsample_merge_tc_blk8: operand sizes must be multiples of 128 bytes");
    {   // r_bytes counts padding between sources; the producer announces a_bytes + r_bytes, so copy sizes must add up
        int rsum = 0;
        for (int r = 0; r < n_resize; ++r) rsum += C * rs_ih[r] * rs_iw[r] * 4;
        PCNN_CHECK_ARG(rsum == roff, "upsample_merge_tc_blk8: resize sources must be multiples of 128 bytes");
    }
    const size_t smem = (size_t)p.w_bytes + 2 * (size_t)p.a_bytes + 2 * (size_t)p.r_bytes + (size_t)NWARP * 32 * C * 4 +
                        (size_t)NWARP * p.rowbuf_floats * 4 + 64;
    PCNN_CHECK_ARG(smem <= 227 * 1024, "upsample_merge_tc_blk8: shared-memory plan too large (%zu bytes)", smem);
    PCNN_CHECK_CUDA(cudaFuncSetAttribute(upsample_merge_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    PCNN_CHECK_CUDA(cudaGetDevice(&dev));
    PCNN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long chunks = (long long)H * ceil_div(W, SEG_W) * ceil_div(B, CH);
    PCNN_CHECK_ARG(chunks < (1ll << 30), "upsample_merge_tc_blk8: too many row segments");
    const int grid = (int)std::min<long long>(chunks, (long long)sms);
    upsample_merge_tc_kernel<<<grid, NTHR, smem, (cudaStream_t)stream>>>(p);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
