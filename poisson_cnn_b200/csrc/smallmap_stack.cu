// Fused convolution stack on tiny 2-D maps (H*W <= 64): the conv + resnet chain of a bottleneck branch whose pooled
// map is 2x2 .. 8x8 (blocks/bottleneck_block.py:36-50 of the reference with downsampling factors 32/64/128 on a
// 256^2 grid).  As separate launches these are 7 latency-bound kernels per branch; here one CTA keeps one sample's
// maps in shared memory for the whole chain (three padded activation buffers) and streams each layer's weights
// (Keras [k][k][Cin][Cout], up to 100 KB) next to them.
//   thread = (4 output channels co4 = tid % 8, pixel group pg = tid / 8 of 16): 4 channels x <= 4 pixels = 16 accumulators;
//   per (input channel, tap) one LDS.128 of weights and four broadcast input reads feed 16 FMAs (the first version,
//   1 channel x 8 pixels per thread, was bound by the load/store unit: 9 shared-memory reads per 8 FMAs).
#include <algorithm>
#include <cstring>

#include "pcnn_common.cuh"
#include "smallmap_stack.h"

namespace pcnn {
namespace sms {

constexpr int MAXL = 24;
constexpr int CMAX = 32;
constexpr int NTHR = 128;
constexpr int MAXPIX = 64;

// asynchronous global -> shared copies (all of a layer's weights in flight at once; the scalar ldg/sts loop this replaces
// exposed one L2 round trip per element)
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 4 : 0;      // src-size 0: zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

struct Params {
    const float* in; float* out;
    const float* kernel[MAXL]; const float* bias[MAXL]; const float* bn_scale[MAXL]; const float* bn_shift[MAXL];
    int k[MAXL], cin[MAXL], cout[MAXL], flags[MAXL];    // flags: 1 = save this layer's input, 2 = add the saved tensor after act/BN
    int n_layers, H, W, act, pad_mode, cin0, padmax, plane;   // plane = (H + 2 padmax) * (W + 2 padmax)
    float pad_value;
};

struct MultiParams { Params prog[MAX_PROGRAMS]; };

// grid (B, programs): blockIdx.y selects the layer program (a bottleneck branch), blockIdx.x the sample
__global__ void __launch_bounds__(NTHR) smallmap_stack_kernel(const __grid_constant__ MultiParams mp) {
    const Params& p = mp.prog[blockIdx.y];
    extern __shared__ __align__(16) float sm[];
    const int H = p.H, W = p.W, pm = p.padmax, Wp = W + 2 * pm, Hp = H + 2 * pm, plane = p.plane;
    float* buf[3] = {sm, sm + CMAX * plane, sm + 2 * CMAX * plane};
    float* s_w = sm + 3 * CMAX * plane;                 // [k*k][cin][32]
    const int tid = threadIdx.x, b = blockIdx.x;
    const int co0 = (tid & 7) * 4, pg = tid >> 3;      // channels co0..co0+3, pixels pg, pg+16, pg+32, pg+48
    const int npix = H * W;

    // (re)write the padding ring of the first c channels (interior already written)
    auto fill_halo = [&](float* a, int c) {
        for (int e = tid; e < c * plane; e += NTHR) {
            const int ch = e / plane, r = e - ch * plane;
            const int y = r / Wp - pm, x = r % Wp - pm;
            if (y >= 0 && y < H && x >= 0 && x < W) continue;
            float v = p.pad_value;
            if (p.pad_mode != PCNN_PAD_CONSTANT)
                v = a[ch * plane + (pad_src_index(y, H, p.pad_mode) + pm) * Wp + pad_src_index(x, W, p.pad_mode) + pm];
            a[e] = v;
        }
    };

    for (int e = tid; e < p.cin0 * npix; e += NTHR) {
        const int ch = e / npix, r = e - ch * npix;
        buf[0][ch * plane + (r / W + pm) * Wp + r % W + pm] = __ldg(p.in + ((long long)b * p.cin0 + ch) * npix + r);
    }
    __syncthreads();
    fill_halo(buf[0], p.cin0);

    // this thread's pixels (offsets inside a padded plane, top-left corner of the window)
    int poff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int pix = pg + 16 * q;
        poff[q] = pix < npix ? (pix / W) * Wp + pix % W : 0;
    }

    int cur = 0, saved = -1;
    for (int l = 0; l < p.n_layers; ++l) {
        const int k = p.k[l], cin = p.cin[l], cout = p.cout[l], pad = k >> 1, taps = k * k;
        if (cout == 32 && (reinterpret_cast<uintptr_t>(p.kernel[l]) & 15) == 0) {
            for (int e = tid; e < taps * cin * 8; e += NTHR) cp_async16(s_w + 4 * e, p.kernel[l] + 4 * (long long)e);
        } else {
            for (int e = tid; e < taps * cin * 32; e += NTHR) {
                const int c = e & 31, r = e >> 5;
                cp_async4(s_w + e, p.kernel[l] + (long long)r * cout + min(c, cout - 1), c < cout);
            }
        }
        cp_async_wait_all();
        if (p.flags[l] & 1) saved = cur;
        int dst = 0;
        while (dst == cur || dst == saved) ++dst;
        __syncthreads();                                 // weights + the input's ring are in place
        const float* a = buf[cur];
        float* o = buf[dst];
        float acc[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[c][q] = 0.f;
        if (co0 < cout && pg < npix) {
            for (int ci = 0; ci < cin; ++ci) {
                const float* base = a + ci * plane + (pm - pad) * Wp + (pm - pad);
                const float* wt = s_w + ci * 32 + co0;
                float part[4][4];                        // blocked summation: one short sum per input channel
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int q = 0; q < 4; ++q) part[c][q] = 0.f;
                for (int dy = 0; dy < k; ++dy)
                    for (int dx = 0; dx < k; ++dx, wt += cin * 32) {
                        const float4 w = *reinterpret_cast<const float4*>(wt);
                        const float* s0 = base + dy * Wp + dx;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float v = s0[poff[q]];
                            part[0][q] = fmaf(v, w.x, part[0][q]); part[1][q] = fmaf(v, w.y, part[1][q]);
                            part[2][q] = fmaf(v, w.z, part[2][q]); part[3][q] = fmaf(v, w.w, part[3][q]);
                        }
                    }
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[c][q] += part[c][q];
            }
            const float* res = (p.flags[l] & 2) ? buf[saved] : nullptr;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int co = co0 + c;
                if (co < cout) {
                    const float bias = p.bias[l] ? __ldg(p.bias[l] + co) : 0.f;
                    const float s = p.bn_scale[l] ? __ldg(p.bn_scale[l] + co) : 1.f, t = p.bn_shift[l] ? __ldg(p.bn_shift[l] + co) : 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (pg + 16 * q < npix) {
                            const int off = co * plane + pm * Wp + pm + poff[q];
                            float v = apply_act(acc[c][q] + bias, p.act);
                            if (p.bn_scale[l]) v = fmaf(v, s, t);
                            if (res) v += res[off];
                            o[off] = v;
                        }
                    }
                }
            }
        }
        __syncthreads();
        fill_halo(o, cout);
        if (p.flags[l] & 2) saved = -1;
        cur = dst;
    }
    __syncthreads();
    const int cl = p.cout[p.n_layers - 1];
    for (int e = tid; e < cl * npix; e += NTHR) {
        const int ch = e / npix, r = e - ch * npix;
        p.out[((long long)b * cl + ch) * npix + r] = buf[cur][ch * plane + (r / W + pm) * Wp + r % W + pm];
    }
    (void)Hp;
}

}  // namespace sms
}  // namespace pcnn

using namespace pcnn;
using namespace pcnn::sms;

static int build_program(const StackDesc& d, Params* out, size_t* smem) {
    const int H = d.H, W = d.W, n_layers = d.n_layers;
    PCNN_CHECK_ARG(d.in && d.out && d.kernels && d.ksize && d.cin && d.cout && d.flags && H > 0 && W > 0, "smallmap_stack_f32: bad argument");
    PCNN_CHECK_ARG(H * W <= MAXPIX, "smallmap_stack_f32: maps of at most %d pixels (got %dx%d)", MAXPIX, H, W);
    PCNN_CHECK_ARG(n_layers >= 1 && n_layers <= MAXL, "smallmap_stack_f32: between 1 and %d layers", MAXL);
    PCNN_CHECK_ARG(d.pad_mode >= PCNN_PAD_CONSTANT && d.pad_mode <= PCNN_PAD_REFLECT, "smallmap_stack_f32: bad pad_mode %d", d.pad_mode);
    PCNN_CHECK_ARG(d.Cin0 >= 1 && d.Cin0 <= CMAX && d.cin[0] == d.Cin0, "smallmap_stack_f32: first layer expects %d input channels", d.Cin0);
    Params& p = *out;
    p.in = d.in; p.out = d.out; p.n_layers = n_layers; p.H = H; p.W = W; p.act = d.act; p.pad_mode = d.pad_mode; p.pad_value = d.pad_value;
    p.cin0 = d.Cin0;
    size_t wmax = 0;
    int depth = 0, padmax = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int k = d.ksize[l];
        PCNN_CHECK_ARG(d.kernels[l] && (k & 1) && k >= 1 && k <= 7, "smallmap_stack_f32: layer %d: odd kernel size <= 7", l);
        PCNN_CHECK_ARG(d.cin[l] >= 1 && d.cin[l] <= CMAX && d.cout[l] >= 1 && d.cout[l] <= CMAX, "smallmap_stack_f32: layer %d: channels must be <= %d", l, CMAX);
        PCNN_CHECK_ARG(l == 0 || d.cin[l] == d.cout[l - 1], "smallmap_stack_f32: layer %d: input channels do not chain", l);
        PCNN_CHECK_ARG((d.bn_scale && d.bn_scale[l]) ? (d.bn_shift && d.bn_shift[l]) : !(d.bn_shift && d.bn_shift[l]), "smallmap_stack_f32: bn_scale/bn_shift must come together");
        if (d.flags[l] & 1) { PCNN_CHECK_ARG(depth == 0, "smallmap_stack_f32: nested saves are not supported"); depth = 1; }
        if (d.flags[l] & 2) { PCNN_CHECK_ARG(depth == 1 || (d.flags[l] & 1), "smallmap_stack_f32: layer %d adds a tensor nobody saved", l); depth = 0; }
        if (d.pad_mode == PCNN_PAD_SYMMETRIC) PCNN_CHECK_ARG(k / 2 <= std::min(H, W), "smallmap_stack_f32: SYMMETRIC pad larger than the map");
        if (d.pad_mode == PCNN_PAD_REFLECT) PCNN_CHECK_ARG(k / 2 < std::min(H, W), "smallmap_stack_f32: REFLECT pad too large for the map");
        p.kernel[l] = d.kernels[l]; p.bias[l] = d.biases ? d.biases[l] : nullptr;
        p.bn_scale[l] = d.bn_scale ? d.bn_scale[l] : nullptr; p.bn_shift[l] = d.bn_shift ? d.bn_shift[l] : nullptr;
        p.k[l] = k; p.cin[l] = d.cin[l]; p.cout[l] = d.cout[l]; p.flags[l] = d.flags[l];
        wmax = std::max(wmax, (size_t)k * k * d.cin[l] * 32);
        padmax = std::max(padmax, k / 2);
    }
    p.padmax = padmax;
    p.plane = (H + 2 * padmax) * (W + 2 * padmax);
    *smem = ((size_t)3 * CMAX * p.plane + wmax) * sizeof(float);
    PCNN_CHECK_ARG(*smem <= 220 * 1024, "smallmap_stack_f32: layer program does not fit in shared memory (%zu B)", *smem);
    return PCNN_OK;
}

int pcnn::sms::smallmap_stack_multi(const StackDesc* descs, int n, int B, cudaStream_t stream) {
    PCNN_CHECK_ARG(descs && n >= 1 && n <= MAX_PROGRAMS && B > 0, "smallmap_stack_f32: between 1 and %d programs per launch", MAX_PROGRAMS);
    MultiParams mp;
    memset(&mp, 0, sizeof(mp));
    size_t smem = 0;
    for (int i = 0; i < n; ++i) {
        size_t s = 0;
        const int rc = build_program(descs[i], &mp.prog[i], &s);
        if (rc != PCNN_OK) return rc;
        smem = std::max(smem, s);
    }
    PCNN_CHECK_CUDA(cudaFuncSetAttribute(smallmap_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smallmap_stack_kernel<<<dim3(B, n), NTHR, smem, stream>>>(mp);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_smallmap_stack_f32(const float* in, float* out, int B, int H, int W, int Cin0, int n_layers,
                                       const float* const* kernels, const float* const* biases,
                                       const float* const* bn_scale, const float* const* bn_shift, const int* ksize,
                                       const int* cin, const int* cout, const int* flags, int act, int pad_mode,
                                       float pad_value, void* stream) {
    const StackDesc d{in, out, H, W, Cin0, n_layers, kernels, biases, bn_scale, bn_shift, ksize, cin, cout, flags, act, pad_mode, pad_value};
    return smallmap_stack_multi(&d, 1, B, (cudaStream_t)stream);
}
