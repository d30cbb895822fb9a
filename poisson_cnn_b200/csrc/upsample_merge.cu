// Fused upsample + merge of the HPNN bottleneck branches (models/Homogeneous_Poisson_NN_Legacy.py:226-233 of the
// reference sums the upsampled branch outputs; blocks/bottleneck_block.py:57-118 upsample each branch with
// deconvupscale (k == stride transpose conv, layers/deconvupscale.py:100-109) or tf.image.resize
// (layers/Upsample.py:57)).  One CTA produces one output row segment of 256 pixels x 32 channels:
//   * every deconv branch: the low-res row that feeds output row Y sits in shared memory; thread groups
//     (column phase tx, 8 channels, 4 low-res pixels) compute all column phases at once on the FP32 FMA
//     pipe (weights as float4 straight from global/L1), stage them phase-major in shared memory, and each
//     of the first 256 threads -- which owns ONE output pixel, all channels -- gathers its phase
//     (conflict-free), applies bias + activation and adds it to its running sums (a [32][256] shared array:
//     registers are left to the 32 accumulators of the compute phase);
//   * every resize branch: the tiny source map sits in shared memory, each thread interpolates its pixel;
//   * the sum (x alpha) is written ONCE, straight into the tensor-core operand layout (BLK8 fp16 + remainder /
//     e4m3 planes) at a channel offset of the concat buffer: the full-resolution fp32 `merged` tensor and its
//     eight read-modify-write passes never exist.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <algorithm>

#include "pcnn_common.cuh"

namespace pcnn {
namespace um {

constexpr int HALO = 7;
constexpr int MAXB = 8;          // branches of each kind
constexpr int SEG_W = 256;       // output pixels per CTA
constexpr int NTHR = 288;        // threads per CTA: 256 pixel owners + one warp so that s = 3 (88 x 3 work items) needs one pass
constexpr float LO_SCALE = 2048.f;

struct Params {
    int n_dc, n_rs;
    const float* dc_in[MAXB];    // [B][Cin][ih][iw]
    const float* dc_w[MAXB];     // Keras deconv kernel [s][s][Cout][Cin]
    const float* dc_b[MAXB];     // [Cout] or null
    int dc_s[MAXB], dc_ih[MAXB], dc_iw[MAXB], dc_pbh[MAXB], dc_pbw[MAXB], dc_act[MAXB], dc_ps[MAXB];
    const float* rs_in[MAXB];    // [B][C][ih][iw]
    const int* rs_iy[MAXB]; const float* rs_wy[MAXB]; const int* rs_ix[MAXB]; const float* rs_wx[MAXB];
    int rs_taps[MAXB], rs_ih[MAXB], rs_iw[MAXB];
    float alpha;
    __half* out; uint8_t* out_lo;
    int mode;                    // precision mode of the destination (1, 2, 3: see pcnn_conv2d_tc)
    int C, H, W, c8_total, plane0;
    int sx_floats;               // size of the input-tile region
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
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// geometry of deconv branch d for this CTA's row segment
struct DcGeom { int s, ih, iw, i, ty, j_begin, segp, pgs; };
__device__ __forceinline__ DcGeom dc_geom(const Params& p, int d, int X0, int Y) {
    DcGeom g;
    g.s = p.dc_s[d]; g.ih = p.dc_ih[d]; g.iw = p.dc_iw[d];
    const int jy = Y + p.dc_pbh[d];
    g.i = jy / g.s; g.ty = jy - g.i * g.s;
    const int xlast = min(X0 + SEG_W - 1, p.W - 1);
    g.j_begin = (X0 + p.dc_pbw[d]) / g.s;
    const int j_end = (xlast + p.dc_pbw[d]) / g.s;
    g.segp = ((j_end - g.j_begin + 1) + 3) & ~3; g.pgs = g.segp >> 2;
    return g;
}

__global__ void __launch_bounds__(NTHR, 2) upsample_merge_kernel(const Params p) {
    extern __shared__ __align__(16) float sm[];
    float* s_sum = sm;                     // [32 channels][256 pixels] running sums
    float* s_tile = sm + 32 * SEG_W;       // 2 x sx_floats: deconv [C][segp] low-res row tile / resize [C][ih][iw] source
    float* s_st = s_tile + 2 * p.sx_floats;   // [s column phases][phase stride] staged deconv results
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int X0 = blockIdx.x * SEG_W, Y = blockIdx.y, b = blockIdx.z;
    const int X = X0 + tid;
    const bool owner = tid < SEG_W && X < p.W;
    const int C = p.C;
    float* my_sum = s_sum + tid;           // owner threads only
    if (tid < SEG_W) {
#pragma unroll
        for (int c = 0; c < 32; ++c) my_sum[c * SEG_W] = 0.f;
    }
    const int nstages = p.n_dc + p.n_rs;

    // asynchronous tile load of stage st (cp.async, one group per stage) + L1 prefetch of this thread's weight rows
    auto issue_load = [&](int st) {
        float* dst = s_tile + (st & 1) * p.sx_floats;
        if (st < p.n_dc) {
            const DcGeom g = dc_geom(p, st, X0, Y);
            const float* inb = p.dc_in[st] + ((long long)b * C * g.ih + g.i) * g.iw + g.j_begin;
            for (int ci = warp; ci < C; ci += NTHR / 32)
                for (int px = lane; px < g.segp; px += 32)
                    cp_async4(dst + ci * g.segp + px, inb + (long long)ci * g.ih * g.iw + px, g.j_begin + px < g.iw);
            if (tid < g.s * 4 * g.pgs) {
                const int tx = tid / (4 * g.pgs);
                const int cg = (tid - tx * 4 * g.pgs) / g.pgs;
                const float* wp = p.dc_w[st] + ((long long)(g.ty * g.s + tx) * C + cg * 8) * C;
                for (int c = 0; c < 8 && cg * 8 + c < C; ++c)
                    for (int k = 0; k < C; k += 8) prefetch_l1(wp + c * C + k);
            }
        } else {
            const int r = st - p.n_dc;
            const int nsrc = C * p.rs_ih[r] * p.rs_iw[r];
            const float* src = p.rs_in[r] + (long long)b * nsrc;
            for (int e = tid; e < nsrc; e += NTHR) cp_async4(dst + e, src + e, true);
        }
    };

    issue_load(0);
    for (int st = 0; st < nstages; ++st) {
        cp_async_commit_wait_all();
        __syncthreads();                   // tile st landed; everyone is done with the other tile buffer and with s_st
        if (st + 1 < nstages) issue_load(st + 1);
        const float* s_x = s_tile + (st & 1) * p.sx_floats;
        if (st < p.n_dc) {
            // ---------------- transpose-conv branch
            const int d = st;
            const DcGeom g = dc_geom(p, d, X0, Y);
            const int s = g.s, segp = g.segp, pgs = g.pgs, ps = p.dc_ps[d];
            const int items = s * 4 * pgs;
            for (int it = tid; it < items; it += NTHR) {
                const int tx = it / (4 * pgs);
                const int rem = it - tx * 4 * pgs;
                const int cg = rem / pgs, pg = rem - cg * pgs;
                if (cg * 8 >= C) continue;
                float acc[8][4];
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[c][q] = 0.f;
                const float* wp = p.dc_w[d] + ((long long)(g.ty * s + tx) * C + cg * 8) * C;
                const float* xp = s_x + pg * 4;
#pragma unroll 1
                for (int ci = 0; ci < C; ci += 4) {
                    float4 x4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) x4[k] = *reinterpret_cast<const float4*>(xp + (ci + k) * segp);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (cg * 8 + c < C) w = __ldg(reinterpret_cast<const float4*>(wp + (long long)c * C + ci));
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
            __syncthreads();
            if (owner) {
                const int jx = X + p.dc_pbw[d];
                const int j = jx / s, tx = jx - j * s;
                const float* gsrc = s_st + tx * ps + (j - g.j_begin);
                const float* bias = p.dc_b[d];
                const int act = p.dc_act[d];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (c < C) {
                        const float v = gsrc[c * segp] + (bias ? __ldg(bias + c) : 0.f);
                        my_sum[c * SEG_W] += (act == PCNN_ACT_LEAKY_RELU) ? fmaxf(v, 0.2f * v) : apply_act(v, act);
                    }
                }
            }
        } else if (owner) {
            // ---------------- resize branch (tiny source in shared memory)
            const int r = st - p.n_dc;
            const int ih = p.rs_ih[r], iw = p.rs_iw[r], taps = p.rs_taps[r];
            int iy[4], ix[4];
            float wy[4], wx[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const bool on = a < taps;
                iy[a] = on ? __ldg(p.rs_iy[r] + Y * taps + a) * iw : 0;
                wy[a] = on ? __ldg(p.rs_wy[r] + Y * taps + a) : 0.f;
                ix[a] = on ? __ldg(p.rs_ix[r] + X * taps + a) : 0;
                wx[a] = on ? __ldg(p.rs_wx[r] + X * taps + a) : 0.f;
            }
#pragma unroll 2
            for (int c = 0; c < C; ++c) {
                const float* src = s_x + c * ih * iw;
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (a < taps) {
                        const float* row = src + iy[a];
                        float rr = 0.f;   // TF interpolates along x first, then along y
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q < taps) rr = fmaf(row[ix[q]], wx[q], rr);
                        acc = fmaf(rr, wy[a], acc);
                    }
                }
                my_sum[c * SEG_W] += acc;
            }
        }
    }

    // ---------------- write-out: one pixel per thread, 16-byte units of the BLK8 layout
    if (!owner) return;
    const int Hp = p.H + 2 * HALO, P = p.W + 2 * HALO;
    const size_t plane_px = (size_t)Hp * P;
    const size_t pix = (size_t)(Y + HALO) * P + (X + HALO);
    const int planes = (C + 7) / 8;
    float sum[32], lo[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) sum[c] = my_sum[c * SEG_W] * p.alpha;
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
                *reinterpret_cast<uint4*>(q) = qv;
                *reinterpret_cast<uint4*>(q + plane_px * 16) = lv;
            }
        }
    }
}

}  // namespace um
}  // namespace pcnn

using namespace pcnn;
using namespace pcnn::um;

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
    PCNN_CHECK_ARG(C >= 4 && C <= 32 && (C % 4) == 0, "upsample_merge_blk8: channels %d must be a multiple of 4, <= 32", C);
    PCNN_CHECK_ARG((c_offset % 16) == 0 && c_offset + C <= ((c_total + 15) / 16) * 16, "upsample_merge_blk8: channel offset must be a multiple of 16 inside the buffer");
    Params p;
    p.n_dc = n_deconv; p.n_rs = n_resize; p.alpha = alpha;
    p.out = (__half*)out; p.out_lo = (uint8_t*)out_lo; p.mode = mode;
    p.C = C; p.H = H; p.W = W; p.c8_total = ((c_total + 15) / 16) * 2; p.plane0 = c_offset / 8;
    int sx = 0, st = 0;
    for (int d = 0; d < n_deconv; ++d) {
        const int s = dc_stride[d], ih = dc_ih[d], iw = dc_iw[d];
        PCNN_CHECK_ARG(dc_in[d] && dc_kernel[d] && s >= 1 && s <= 64, "upsample_merge_blk8: deconv branch %d: bad argument", d);
        PCNN_CHECK_ARG((reinterpret_cast<uintptr_t>(dc_kernel[d]) % 16) == 0, "upsample_merge_blk8: deconv kernel %d must be 16-byte aligned", d);
        PCNN_CHECK_ARG(ceil_div(H, s) == ih && ceil_div(W, s) == iw,
                       "upsample_merge_blk8: output (%d,%d) inconsistent with input (%d,%d) at stride %d (TF raises)", H, W, ih, iw, s);
        p.dc_in[d] = dc_in[d]; p.dc_w[d] = dc_kernel[d]; p.dc_b[d] = dc_bias ? dc_bias[d] : nullptr;
        p.dc_s[d] = s; p.dc_ih[d] = ih; p.dc_iw[d] = iw; p.dc_act[d] = dc_act[d];
        p.dc_pbh[d] = std::max((ih - 1) * s + s - H, 0) / 2;
        p.dc_pbw[d] = std::max((iw - 1) * s + s - W, 0) / 2;
        const int segp_max = (((SEG_W - 1) / s + 2) + 3) & ~3;           // low-res pixels touching one 256-pixel segment
        int ps = 32 * segp_max + (32 + s - 1) / s;                        // phases land ~32/s banks apart
        ps = (ps + 3) & ~3;
        p.dc_ps[d] = ps;
        sx = std::max(sx, 32 * segp_max);
        st = std::max(st, s * ps);
    }
    for (int r = 0; r < n_resize; ++r) {
        PCNN_CHECK_ARG(rs_in[r] && rs_iy[r] && rs_wy[r] && rs_ix[r] && rs_wx[r] && rs_taps[r] >= 1 && rs_taps[r] <= 4 && rs_ih[r] > 0 && rs_iw[r] > 0,
                       "upsample_merge_blk8: resize branch %d: bad argument", r);
        PCNN_CHECK_ARG((long long)C * rs_ih[r] * rs_iw[r] <= 8192, "upsample_merge_blk8: resize source %dx%d too large for the fused kernel", rs_ih[r], rs_iw[r]);
        p.rs_in[r] = rs_in[r]; p.rs_iy[r] = rs_iy[r]; p.rs_wy[r] = rs_wy[r]; p.rs_ix[r] = rs_ix[r]; p.rs_wx[r] = rs_wx[r];
        p.rs_taps[r] = rs_taps[r]; p.rs_ih[r] = rs_ih[r]; p.rs_iw[r] = rs_iw[r];
        sx = std::max(sx, C * rs_ih[r] * rs_iw[r]);
    }
    sx = (sx + 3) & ~3;
    p.sx_floats = sx;
    const size_t smem = (size_t)(32 * SEG_W + 2 * sx + st) * sizeof(float);
    PCNN_CHECK_ARG(smem <= 200 * 1024, "upsample_merge_blk8: shared-memory plan too large (%zu bytes)", smem);
    if (smem > 48 * 1024)
        PCNN_CHECK_CUDA(cudaFuncSetAttribute(upsample_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upsample_merge_kernel<<<dim3(ceil_div(W, SEG_W), H, B), NTHR, smem, (cudaStream_t)stream>>>(p);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
