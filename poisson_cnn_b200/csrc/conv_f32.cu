// Strict-FP32 padded direct convolution (CUDA-core FFMA) with fused epilogue.
// This is the 1e-5 "strict" path; the tensor-core path lives in conv_tc.cu.
//
// Tile: one CTA computes an 8 x 64 pixel tile for all output channels of one sample.
//   threadIdx.x in [0,64): row r = tid/8, column lane cx = tid%8; the thread owns the 8 pixels
//   x = cx + 8*p (p=0..7) of row r  -> shared-memory reads of the input tile are conflict-free
//   (8 consecutive floats per row, 4 rows per warp at pitch == 8 mod 32).
//   threadIdx.y = group of 8 output channels (a warp reads its 8 weights as two broadcast LDS.128).
// Input channels are streamed through shared memory in chunks; the halo (tf.pad CONSTANT /
// SYMMETRIC / REFLECT) is resolved while the tile is loaded, so no padded copy ever exists.
#include <algorithm>

#include "pcnn_common.cuh"

namespace pcnn {

constexpr int TILE_H = 8;
constexpr int TILE_W = 64;
constexpr int PX = 8;    // pixels per thread
constexpr int CO = 8;    // output channels per thread

struct ConvF32Params {
    const float* in; const float* kernel; const float* bias; const float* bn_scale;
    const float* bn_shift; const float* residual; const float* out_scale; float* out;
    int B, Cin, Cout, H, W, kh, kw, pad_mode, act;
    float pad_value;
    long long in_bstride, out_bstride, res_bstride;
    // channel / row strides of the tiled kernel (dense NCHW: H*W and W; a batch of 1-D signals run as the rows
    // of one image: W and the sample stride)
    long long in_cs, in_rs, out_cs, out_rs, res_cs, res_rs;
    int ci_chunk, pitch, tile_rows, cop;   // cop = padded Cout (multiple of 8)
    // Blocked summation: every flush_ci input channels (~256 product terms) the register accumulators are added to
    // per-thread totals in shared memory and restart from zero, so a K = kh*kw*Cin = 7200-term dot product is a sum of
    // ~30 short sums instead of one long sequential FP32 sum (rounding error ~sqrt(256)+sqrt(30) instead of ~sqrt(7200)
    // ulps: the strict mode's distance to the float64 oracle at 256x256 went from 1.4e-5 to below the 1e-5 budget).
    // 0 = a single sequential sum (short dot products).
    int flush_ci;
};

__global__ void __launch_bounds__(256) conv2d_f32_kernel(const ConvF32Params p) {
    extern __shared__ float smem[];
    const int tile_elems = p.tile_rows * p.pitch;
    float* s_in = smem;                                   // [ci_chunk][tile_rows][pitch]
    float* s_w = smem + (size_t)p.ci_chunk * tile_elems;   // [ci_chunk][kh][kw][cop]
    float* s_tot = s_w + (size_t)p.ci_chunk * p.kh * p.kw * p.cop;   // [PX*CO][threads] running totals (flush_ci > 0)

    const int tid = threadIdx.x, cg = threadIdx.y;
    const int nthreads = blockDim.x * blockDim.y;
    const int flat = cg * blockDim.x + tid;
    const int r = tid >> 3, cx = tid & 7;
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * TILE_H, x0 = blockIdx.x * TILE_W;
    const int pad_t = p.kh / 2, pad_l = p.kw / 2;
    const int tile_cols = TILE_W + p.kw - 1;
    const float* inb = p.in + (long long)b * p.in_bstride;

    float acc[PX][CO];
#pragma unroll
    for (int i = 0; i < PX; ++i)
#pragma unroll
        for (int j = 0; j < CO; ++j) acc[i][j] = 0.f;

    const int wk = p.kh * p.kw * p.cop;   // smem weights per input channel
    int since_flush = 0;
    if (p.flush_ci) {
#pragma unroll
        for (int o = 0; o < PX * CO; ++o) s_tot[o * nthreads + flat] = 0.f;     // private to this thread: no barrier needed
    }

    for (int c0 = 0; c0 < p.Cin; c0 += p.ci_chunk) {
        const int nci = min(p.ci_chunk, p.Cin - c0);
        __syncthreads();
        // ---- input tile with halo ----
        for (int idx = flat; idx < nci * tile_elems; idx += nthreads) {
            const int ci = idx / tile_elems;
            const int rem = idx - ci * tile_elems;
            const int tr = rem / p.pitch, tc = rem - tr * p.pitch;
            float v = 0.f;
            if (tc < tile_cols) {
                const int gy = y0 + tr - pad_t, gx = x0 + tc - pad_l;
                const bool inside = (gy >= 0) & (gy < p.H) & (gx >= 0) & (gx < p.W);
                if (inside || p.pad_mode != PCNN_PAD_CONSTANT) {
                    const int sy = pad_src_index(gy, p.H, p.pad_mode);
                    const int sx = pad_src_index(gx, p.W, p.pad_mode);
                    v = __ldg(inb + (long long)(c0 + ci) * p.in_cs + (long long)sy * p.in_rs + sx);
                } else {
                    v = p.pad_value;
                }
            }
            s_in[idx] = v;
        }
        // ---- weights: global [kh,kw,Cin,Cout] -> smem [ci][kh*kw][cop] (zero padded) ----
        for (int idx = flat; idx < nci * wk; idx += nthreads) {
            const int ci = idx / wk;
            const int rem = idx - ci * wk;
            const int tap = rem / p.cop, co = rem - tap * p.cop;
            float v = 0.f;
            if (co < p.Cout) v = __ldg(p.kernel + ((long long)tap * p.Cin + (c0 + ci)) * p.Cout + co);
            s_w[idx] = v;
        }
        __syncthreads();

        for (int ci = 0; ci < nci; ++ci) {
            const float* tin = s_in + ci * tile_elems + r * p.pitch + cx;
            const float* tw = s_w + ci * wk + cg * CO;
            for (int dy = 0; dy < p.kh; ++dy) {
                const float* row = tin + dy * p.pitch;
                const float* wrow = tw + dy * p.kw * p.cop;
#pragma unroll 3
                for (int dx = 0; dx < p.kw; ++dx) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wrow + dx * p.cop);
                    const float4 w1 = *reinterpret_cast<const float4*>(wrow + dx * p.cop + 4);
                    float v[PX];
#pragma unroll
                    for (int i = 0; i < PX; ++i) v[i] = row[dx + 8 * i];
#pragma unroll
                    for (int i = 0; i < PX; ++i) {
                        acc[i][0] = fmaf(v[i], w0.x, acc[i][0]);
                        acc[i][1] = fmaf(v[i], w0.y, acc[i][1]);
                        acc[i][2] = fmaf(v[i], w0.z, acc[i][2]);
                        acc[i][3] = fmaf(v[i], w0.w, acc[i][3]);
                        acc[i][4] = fmaf(v[i], w1.x, acc[i][4]);
                        acc[i][5] = fmaf(v[i], w1.y, acc[i][5]);
                        acc[i][6] = fmaf(v[i], w1.z, acc[i][6]);
                        acc[i][7] = fmaf(v[i], w1.w, acc[i][7]);
                    }
                }
            }
            if (p.flush_ci && ++since_flush == p.flush_ci) {
                since_flush = 0;
#pragma unroll
                for (int i = 0; i < PX; ++i)
#pragma unroll
                    for (int j = 0; j < CO; ++j) {
                        s_tot[(i * CO + j) * nthreads + flat] += acc[i][j];
                        acc[i][j] = 0.f;
                    }
            }
        }
    }
    if (p.flush_ci) {
#pragma unroll
        for (int i = 0; i < PX; ++i)
#pragma unroll
            for (int j = 0; j < CO; ++j) acc[i][j] += s_tot[(i * CO + j) * nthreads + flat];
    }

    // ---- epilogue: bias -> activation -> BN affine -> residual -> per-(b,c) scale ----
    const int gy = y0 + r;
    if (gy >= p.H) return;
    float* outb = p.out + (long long)b * p.out_bstride;
    const float* resb = p.residual ? p.residual + (long long)b * p.res_bstride : nullptr;
#pragma unroll
    for (int j = 0; j < CO; ++j) {
        const int co = cg * CO + j;
        if (co >= p.Cout) break;
        const float bias = p.bias ? __ldg(p.bias + co) : 0.f;
        const float bs = p.bn_scale ? __ldg(p.bn_scale + co) : 1.f;
        const float bt = p.bn_shift ? __ldg(p.bn_shift + co) : 0.f;
        const float os = p.out_scale ? __ldg(p.out_scale + (long long)b * p.Cout + co) : 1.f;
        const long long base = (long long)co * p.out_cs + (long long)gy * p.out_rs;
        const long long rbase = (long long)co * p.res_cs + (long long)gy * p.res_rs;
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            const int gx = x0 + cx + 8 * i;
            if (gx < p.W) {
                float v = apply_act(acc[i][j] + bias, p.act);
                if (p.bn_scale) v = fmaf(v, bs, bt);
                if (resb) v += __ldg(resb + rbase + gx);
                if (p.out_scale) v *= os;
                outb[base + gx] = v;
            }
        }
    }
}

// Small-map variant (H*W <= 64: the 2x2 .. 8x8 maps of the ds = 32/64/128 branches and the Scaling
// tail): one CTA per sample keeps the whole padded input in shared memory; thread = (output channel,
// pixel group), weights read coalesced across channels.  The 8x64-pixel tile kernel above would spend
// >85 % of its lanes on pixels that do not exist.
__global__ void __launch_bounds__(256) conv2d_small_f32_kernel(const ConvF32Params p) {
    extern __shared__ float smem[];
    const int b = blockIdx.x;
    const int co0 = blockIdx.y * 8;             // 8 output channels per CTA (4 CTAs per sample at Cout = 32)
    const int ph = p.H + p.kh - 1, pw = p.W + p.kw - 1;
    const int pad_t = p.kh / 2, pad_l = p.kw / 2;
    const int taps = p.kh * p.kw;
    float* s_in = smem;                          // [Cin][ph][pw]
    float* s_w = smem + p.Cin * ph * pw;         // [Cin][taps][8]
    const float* inb = p.in + (long long)b * p.in_bstride;
    for (int idx = threadIdx.x; idx < p.Cin * ph * pw; idx += blockDim.x) {
        const int ci = idx / (ph * pw);
        const int rem = idx - ci * ph * pw;
        const int yy = rem / pw - pad_t, xx = rem % pw - pad_l;
        const bool inside = (yy >= 0) & (yy < p.H) & (xx >= 0) & (xx < p.W);
        float v = p.pad_value;
        if (inside || p.pad_mode != PCNN_PAD_CONSTANT)
            v = __ldg(inb + ((long long)ci * p.H + pad_src_index(yy, p.H, p.pad_mode)) * p.W + pad_src_index(xx, p.W, p.pad_mode));
        s_in[idx] = v;
    }
    for (int idx = threadIdx.x; idx < p.Cin * taps * 8; idx += blockDim.x) {
        const int c = idx & 7;
        const int t = (idx >> 3) % taps, ci = (idx >> 3) / taps;
        s_w[idx] = (co0 + c < p.Cout) ? __ldg(p.kernel + ((long long)t * p.Cin + ci) * p.Cout + co0 + c) : 0.f;
    }
    __syncthreads();
    const int npix = p.H * p.W;
    // thread = (channel c of 8, pixel slot of 32): pixels pslot and pslot + 32
    const int c = threadIdx.x & 7, pslot = threadIdx.x >> 3;
    const int co = co0 + c;
    float acc[2] = {0.f, 0.f};
    int poff[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int pidx = pslot + 32 * i;
        poff[i] = pidx < npix ? (pidx / p.W) * pw + (pidx % p.W) : 0;
    }
    for (int ci = 0; ci < p.Cin; ++ci) {
        const float* sin = s_in + ci * ph * pw;
        const float* sw = s_w + ci * taps * 8 + c;
        float part[2] = {0.f, 0.f};              // blocked summation: one short sum per input channel
        for (int dy = 0; dy < p.kh; ++dy)
            for (int dx = 0; dx < p.kw; ++dx) {
                const float w = sw[(dy * p.kw + dx) * 8];
                const float* s0 = sin + dy * pw + dx;
                part[0] = fmaf(s0[poff[0]], w, part[0]);
                part[1] = fmaf(s0[poff[1]], w, part[1]);
            }
        acc[0] += part[0];
        acc[1] += part[1];
    }
    if (co >= p.Cout) return;
    const float bias = p.bias ? __ldg(p.bias + co) : 0.f;
    const float bs = p.bn_scale ? __ldg(p.bn_scale + co) : 1.f, bt = p.bn_shift ? __ldg(p.bn_shift + co) : 0.f;
    const float os = p.out_scale ? __ldg(p.out_scale + (long long)b * p.Cout + co) : 1.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int pidx = pslot + 32 * i;
        if (pidx >= npix) break;
        float v = apply_act(acc[i] + bias, p.act);
        if (p.bn_scale) v = fmaf(v, bs, bt);
        const long long off = (long long)co * npix + pidx;
        if (p.residual) v += __ldg(p.residual + (long long)b * p.res_bstride + off);
        if (p.out_scale) v *= os;
        p.out[(long long)b * p.out_bstride + off] = v;
    }
}

}  // namespace pcnn

extern "C" int pcnn_conv2d_f32(const float* in, const float* kernel, const float* bias,
                               const float* bn_scale, const float* bn_shift, const float* residual,
                               const float* out_scale, float* out, int B, int Cin, int Cout, int H,
                               int W, int kh, int kw, int pad_mode, float pad_value, int act,
                               int64_t in_bstride, int64_t out_bstride, int64_t res_bstride,
                               void* stream) {
    using namespace pcnn;
    PCNN_CHECK_ARG(in && kernel && out, "conv2d_f32: null pointer");
    PCNN_CHECK_ARG(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && kh > 0 && kw > 0,
                   "conv2d_f32: non-positive dimension");
    PCNN_CHECK_ARG(B <= 65535, "conv2d_f32: batch %d exceeds grid.z limit", B);
    PCNN_CHECK_ARG((bn_scale == nullptr) == (bn_shift == nullptr), "conv2d_f32: bn_scale/bn_shift must come together");
    PCNN_CHECK_ARG(pad_mode >= PCNN_PAD_CONSTANT && pad_mode <= PCNN_PAD_REFLECT, "conv2d_f32: bad pad_mode %d", pad_mode);
    if (pad_mode == PCNN_PAD_SYMMETRIC)
        PCNN_CHECK_ARG(kh / 2 <= H && kw / 2 <= W, "conv2d_f32: SYMMETRIC pad (%d,%d) larger than input (%d,%d)", kh / 2, kw / 2, H, W);
    if (pad_mode == PCNN_PAD_REFLECT)
        PCNN_CHECK_ARG(kh / 2 < H && kw / 2 < W, "conv2d_f32: REFLECT pad too large for input");

    ConvF32Params p;
    p.in = in; p.kernel = kernel; p.bias = bias; p.bn_scale = bn_scale; p.bn_shift = bn_shift;
    p.residual = residual; p.out_scale = out_scale; p.out = out;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.kh = kh; p.kw = kw;
    p.pad_mode = pad_mode; p.act = act; p.pad_value = pad_value;
    p.in_bstride = in_bstride; p.out_bstride = out_bstride; p.res_bstride = res_bstride;
    const int ngroups = ceil_div(Cout, CO);
    PCNN_CHECK_ARG(ngroups <= 4, "conv2d_f32: Cout %d > 32 not supported", Cout);
    {
        const size_t small_smem = ((size_t)Cin * (H + kh - 1) * (W + kw - 1) + (size_t)Cin * kh * kw * 8) * sizeof(float);
        if (H * W <= 64 && small_smem <= 160 * 1024) {
            p.cop = 0; p.tile_rows = 0; p.pitch = 0; p.ci_chunk = 0;
            if (small_smem > 48 * 1024)
                PCNN_CHECK_CUDA(cudaFuncSetAttribute(conv2d_small_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            conv2d_small_f32_kernel<<<dim3(B, ceil_div(Cout, 8)), 256, small_smem, (cudaStream_t)stream>>>(p);
            PCNN_CHECK_LAUNCH();
            return PCNN_OK;
        }
    }
    p.in_cs = p.out_cs = p.res_cs = (long long)H * W;
    p.in_rs = p.out_rs = p.res_rs = W;
    int gridB = B;
    if (H == 1 && kh == 1 && B > 1 && !out_scale) {
        // a batch of 1-D signals (the DBCNN boundary stack): the samples become the rows of ONE image, so the
        // 8-row tiles are full instead of 1/8 occupied; kh == 1 keeps the rows independent
        p.H = B; p.B = 1; gridB = 1;
        p.in_cs = p.out_cs = p.res_cs = W;
        p.in_rs = in_bstride; p.out_rs = out_bstride; p.res_rs = res_bstride;
    }
    p.cop = ngroups * CO;
    p.tile_rows = TILE_H + kh - 1;
    int cols = TILE_W + kw - 1;
    p.pitch = cols + ((8 - (cols % 32)) % 32 + 32) % 32;   // pitch == 8 (mod 32)
    const size_t per_ci = ((size_t)p.tile_rows * p.pitch + (size_t)kh * kw * p.cop) * sizeof(float);
    // blocked summation for long dot products (see ConvF32Params::flush_ci): 64 KB of per-thread totals at 256 threads
    const int taps = kh * kw;
    p.flush_ci = ((long long)taps * Cin > 512) ? std::max(1, 256 / taps) : 0;
    const size_t tot_bytes = p.flush_ci ? (size_t)PX * CO * 64 * ngroups * sizeof(float) : 0;
    const size_t budget = (p.flush_ci ? 48 : 96) * 1024;
    int chunk = (int)(budget / per_ci);
    if (chunk < 1) chunk = 1;
    if (chunk > Cin) chunk = Cin;
    if (chunk > 8) chunk = 8;
    p.ci_chunk = chunk;
    const size_t smem = per_ci * chunk + tot_bytes;
    PCNN_CHECK_ARG(smem <= 200 * 1024, "conv2d_f32: kernel %dx%d needs %zu B of shared memory", kh, kw, smem);
    // the attribute belongs to the (device, context) the launch goes to: set it on every large launch (cheap) instead of
    // caching one process-wide flag, which broke the second GPU of a multi-device process
    if (smem > 48 * 1024)
        PCNN_CHECK_CUDA(cudaFuncSetAttribute(conv2d_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 block(64, ngroups, 1);
    dim3 grid(ceil_div(W, TILE_W), ceil_div(p.H, TILE_H), gridB);
    PCNN_CHECK_ARG(grid.y <= 65535, "conv2d_f32: H too large");
    conv2d_f32_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(p);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
