// Fused 1-D convolution stack of the Dirichlet boundary network (models/Dirichlet_BC_NN_Legacy.py:47-64,136-141 of the
// reference: per stage a Conv1D (+BatchNorm) followed by a resnet of three Conv1Ds, 8 stages = 32 convolutions on a
// [<=27 channels, n] signal).  As separate launches these are 32 latency-bound kernels of ~65 us each; here one CTA
// keeps one boundary signal in shared memory for the whole stack: three activation buffers (input, output, the
// resnet's saved input) with a materialised padding halo, the current layer's weights next to them.
//   thread = (group of 4 consecutive positions, channel group cg): channels cg, cg+4, ..., <= 7 of them, 28 FP32
//   accumulators; per (input channel, tap) one new input value slides into a 4-wide register window and 7 weights are
//   read as warp-wide broadcasts.
#include <algorithm>

#include "pcnn_common.cuh"

namespace pcnn {
namespace bs {

constexpr int MAXL = 48;      // layers in a stack
constexpr int PADMAX = 9;     // kernel sizes up to 19
constexpr int CMAX = 28;      // channels (4 channel groups x 7)
constexpr int NTHR = 256;

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
    int n_layers, n, act, pad_mode, cin0, pitch;
    float pad_value;
};

__global__ void __launch_bounds__(NTHR) boundary_stack_kernel(const Params p) {
    extern __shared__ __align__(16) float sm[];
    const int pitch = p.pitch, n = p.n;
    float* buf[3] = {sm, sm + CMAX * pitch, sm + 2 * CMAX * pitch};
    float* s_w = sm + 3 * CMAX * pitch;                 // [k][cin][32]
    const int tid = threadIdx.x, b = blockIdx.x;

    // write the padding halo of the first c channels of a buffer (tf.pad CONSTANT / SYMMETRIC / REFLECT, any width <= PADMAX)
    auto fill_halo = [&](float* a, int c) {
        for (int e = tid; e < c * 2 * PADMAX; e += NTHR) {
            const int ch = e / (2 * PADMAX), j = e - ch * 2 * PADMAX;
            const int x = j < PADMAX ? j - PADMAX : n + (j - PADMAX);
            float v = p.pad_value;
            if (p.pad_mode != PCNN_PAD_CONSTANT) v = a[ch * pitch + PADMAX + pad_src_index(x, n, p.pad_mode)];
            a[ch * pitch + PADMAX + x] = v;
        }
    };

    for (int e = tid; e < p.cin0 * n; e += NTHR) {
        const int ch = e / n, x = e - ch * n;
        buf[0][ch * pitch + PADMAX + x] = __ldg(p.in + ((long long)b * p.cin0 + ch) * n + x);
    }
    __syncthreads();
    fill_halo(buf[0], p.cin0);

    int cur = 0, saved = -1;
    const int pgs = (n + 3) >> 2;
    for (int l = 0; l < p.n_layers; ++l) {
        const int k = p.k[l], cin = p.cin[l], cout = p.cout[l], pad = k >> 1;
        // weights: Keras Conv1D [k][cin][cout] -> smem [k][cin][32]
        for (int e = tid; e < k * cin * 32; e += NTHR) {
            const int co = e & 31, r = e >> 5;
            cp_async4(s_w + e, p.kernel[l] + (long long)r * cout + min(co, cout - 1), co < cout);
        }
        cp_async_wait_all();
        if (p.flags[l] & 1) saved = cur;
        int dst = 0;
        while (dst == cur || dst == saved) ++dst;
        __syncthreads();                                 // weights + the input's halo are in place
        const float* a = buf[cur];
        float* o = buf[dst];
        const float* res = (p.flags[l] & 2) ? buf[saved] : nullptr;
        for (int it = tid; it < 4 * pgs; it += NTHR) {
            const int cg = it / pgs, pg = it - cg * pgs;
            if (cg >= cout) continue;
            const int x0 = pg * 4;
            float acc[7][4];
#pragma unroll
            for (int j = 0; j < 7; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
            for (int ci = 0; ci < cin; ++ci) {
                const float* row = a + ci * pitch + PADMAX + x0 - pad;
                float w0 = row[0], w1 = row[1], w2 = row[2], w3;
                const float* wt = s_w + ci * 32 + cg;
                for (int t = 0; t < k; ++t, wt += cin * 32) {
                    w3 = row[t + 3];
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const float w = wt[4 * j];
                        acc[j][0] = fmaf(w0, w, acc[j][0]); acc[j][1] = fmaf(w1, w, acc[j][1]);
                        acc[j][2] = fmaf(w2, w, acc[j][2]); acc[j][3] = fmaf(w3, w, acc[j][3]);
                    }
                    w0 = w1; w1 = w2; w2 = w3;
                }
            }
            // epilogue: bias -> activation -> BN affine -> + saved resnet input
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int co = cg + 4 * j;
                if (co < cout) {
                    const float bias = p.bias[l] ? __ldg(p.bias[l] + co) : 0.f;
                    const float s = p.bn_scale[l] ? __ldg(p.bn_scale[l] + co) : 1.f, t = p.bn_shift[l] ? __ldg(p.bn_shift[l] + co) : 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (x0 + q < n) {
                            float v = apply_act(acc[j][q] + bias, p.act);
                            if (p.bn_scale[l]) v = fmaf(v, s, t);
                            if (res) v += res[co * pitch + PADMAX + x0 + q];
                            o[co * pitch + PADMAX + x0 + q] = v;
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
    for (int e = tid; e < cl * n; e += NTHR) {
        const int ch = e / n, x = e - ch * n;
        p.out[((long long)b * cl + ch) * n + x] = buf[cur][ch * pitch + PADMAX + x];
    }
}

}  // namespace bs
}  // namespace pcnn

using namespace pcnn;
using namespace pcnn::bs;

extern "C" int pcnn_boundary_stack_f32(const float* in, float* out, int B, int n, int Cin0, int n_layers,
                                       const float* const* kernels, const float* const* biases,
                                       const float* const* bn_scale, const float* const* bn_shift, const int* ksize,
                                       const int* cin, const int* cout, const int* flags, int act, int pad_mode,
                                       float pad_value, void* stream) {
    PCNN_CHECK_ARG(in && out && kernels && ksize && cin && cout && flags && B > 0 && n > 0, "boundary_stack_f32: bad argument");
    PCNN_CHECK_ARG(n_layers >= 1 && n_layers <= MAXL, "boundary_stack_f32: between 1 and %d layers", MAXL);
    PCNN_CHECK_ARG(pad_mode >= PCNN_PAD_CONSTANT && pad_mode <= PCNN_PAD_REFLECT, "boundary_stack_f32: bad pad_mode %d", pad_mode);
    PCNN_CHECK_ARG(Cin0 >= 1 && Cin0 <= CMAX && cin[0] == Cin0, "boundary_stack_f32: first layer expects %d input channels", Cin0);
    Params p;
    p.in = in; p.out = out; p.n_layers = n_layers; p.n = n; p.act = act; p.pad_mode = pad_mode; p.pad_value = pad_value; p.cin0 = Cin0;
    p.pitch = ((n + 2 * PADMAX + 3) / 4) * 4;
    size_t wmax = 0;
    int depth = 0;
    for (int l = 0; l < n_layers; ++l) {
        PCNN_CHECK_ARG(kernels[l] && (ksize[l] & 1) && ksize[l] >= 1 && ksize[l] <= 2 * PADMAX + 1, "boundary_stack_f32: layer %d: odd kernel size <= %d", l, 2 * PADMAX + 1);
        PCNN_CHECK_ARG(cin[l] >= 1 && cin[l] <= CMAX && cout[l] >= 1 && cout[l] <= CMAX, "boundary_stack_f32: layer %d: channels must be <= %d", l, CMAX);
        PCNN_CHECK_ARG(l == 0 || cin[l] == cout[l - 1], "boundary_stack_f32: layer %d: input channels do not chain", l);
        PCNN_CHECK_ARG((bn_scale && bn_scale[l]) ? (bn_shift && bn_shift[l]) : !(bn_shift && bn_shift[l]), "boundary_stack_f32: bn_scale/bn_shift must come together");
        if (flags[l] & 1) { PCNN_CHECK_ARG(depth == 0, "boundary_stack_f32: nested saves are not supported"); depth = 1; }
        if (flags[l] & 2) { PCNN_CHECK_ARG(depth == 1 || (flags[l] & 1), "boundary_stack_f32: layer %d adds a tensor nobody saved", l); depth = 0; }
        if (pad_mode == PCNN_PAD_SYMMETRIC) PCNN_CHECK_ARG(ksize[l] / 2 <= n, "boundary_stack_f32: SYMMETRIC pad larger than the signal");
        if (pad_mode == PCNN_PAD_REFLECT) PCNN_CHECK_ARG(ksize[l] / 2 < n, "boundary_stack_f32: REFLECT pad too large for the signal");
        p.kernel[l] = kernels[l]; p.bias[l] = biases ? biases[l] : nullptr;
        p.bn_scale[l] = bn_scale ? bn_scale[l] : nullptr; p.bn_shift[l] = bn_shift ? bn_shift[l] : nullptr;
        p.k[l] = ksize[l]; p.cin[l] = cin[l]; p.cout[l] = cout[l]; p.flags[l] = flags[l];
        wmax = std::max(wmax, (size_t)ksize[l] * cin[l] * 32);
    }
    const size_t smem = ((size_t)3 * CMAX * p.pitch + wmax) * sizeof(float);
    PCNN_CHECK_ARG(smem <= 220 * 1024, "boundary_stack_f32: signal of %d samples does not fit in shared memory (%zu B)", n, smem);
    PCNN_CHECK_CUDA(cudaFuncSetAttribute(boundary_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    boundary_stack_kernel<<<B, NTHR, smem, (cudaStream_t)stream>>>(p);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
