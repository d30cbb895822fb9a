// Memory-bound kernels of the hot path: SAME average pooling, transpose-conv / resize upsampling
// with fused merge accumulation, spatial pyramid pooling, Dense, per-sample max-normalisation and
// the model glue (input assembly, sinh-mode expansion, boundary ring, oriented 5-way merge).
// All fp32, NCHW, coalesced along W; reductions use warp shuffles.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cmath>

#include "pcnn_common.cuh"

namespace pcnn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------ avg pool (SAME, pool=stride=s)
// grid (ceil(ow/128), oh, B*C): no per-element div/mod.  Small windows: one thread per output.
__global__ void __launch_bounds__(128) avgpool_thread_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                             int H, int W, int oh, int ow, int s, int pt, int pl,
                                                             long long in_bstride) {
    const int ox = blockIdx.x * 128 + threadIdx.x, oy = blockIdx.y;
    if (ox >= ow) return;
    const int b = blockIdx.z / C, c = blockIdx.z - b * C;
    const int ys = max(oy * s - pt, 0), ye = min(oy * s - pt + s, H);
    const int xs = max(ox * s - pl, 0), xe = min(ox * s - pl + s, W);
    const float* src = in + (long long)b * in_bstride + (long long)c * H * W;
    float acc = 0.f;
    for (int y = ys; y < ye; ++y)
        for (int x = xs; x < xe; ++x) acc += __ldg(src + (long long)y * W + x);
    out[((long long)blockIdx.z * oh + oy) * ow + ox] = acc / (float)((ye - ys) * (xe - xs));
}

// Large windows (s >= 8): a CTA owns one (sample, channel) plane and a band of output rows.  Per output row the
// window rows are first reduced to per-column sums (one column per thread, fully coalesced reads, the row loop
// unrolled so that several loads are in flight), then each warp reduces whole windows out of shared memory.
// grid (ceil(oh/rows_per_cta), B*C), 256 threads, W floats of shared memory.
__global__ void __launch_bounds__(256) avgpool_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                           int H, int W, int oh, int ow, int s, int pt, int pl,
                                                           long long in_bstride, int rows_per_cta) {
    extern __shared__ float colsum[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y / C, c = blockIdx.y - b * C;
    const float* src = in + (long long)b * in_bstride + (long long)c * H * W;
    for (int r = 0; r < rows_per_cta; ++r) {
        const int oy = blockIdx.x * rows_per_cta + r;
        if (oy >= oh) break;                                  // block-uniform
        const int ys = max(oy * s - pt, 0), ye = min(oy * s - pt + s, H);
        for (int x = threadIdx.x; x < W; x += 256) {
            const float* col = src + (long long)ys * W + x;
            float acc = 0.f;
            int y = ys;
            for (; y + 8 <= ye; y += 8, col += 8ll * W) {
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldg(col + (long long)k * W);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += v[k];
            }
            for (; y < ye; ++y, col += W) acc += __ldg(col);
            colsum[x] = acc;
        }
        __syncthreads();
        for (int ox = warp; ox < ow; ox += 8) {
            const int xs = max(ox * s - pl, 0), xe = min(ox * s - pl + s, W);
            float acc = 0.f;
            for (int x = xs + lane; x < xe; x += 32) acc += colsum[x];
            acc = warp_sum(acc);
            if (lane == 0) out[((long long)blockIdx.y * oh + oy) * ow + ox] = acc / (float)((ye - ys) * (xe - xs));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ transpose conv (SAME)
// one thread per output element; taps enumerated directly: t == (Y+pb) mod s (+ m*s)
__global__ void deconv_same_kernel(const float* __restrict__ in, const float* __restrict__ kernel,
                                   const float* __restrict__ bias, float* __restrict__ out, int Cin,
                                   int Cout, int ih, int iw, int oh, int ow, int kh, int kw, int s,
                                   int pbh, int pbw, int act, float alpha, int accumulate,
                                   long long out_bstride, long long total) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int X = idx % ow;
        long long t = idx / ow;
        const int Y = t % oh; t /= oh;
        const int co = t % Cout;
        const int b = t / Cout;
        float acc = bias ? __ldg(bias + co) : 0.f;
        const float* inb = in + (long long)b * Cin * ih * iw;
        for (int ty = (Y + pbh) % s; ty < kh; ty += s) {
            const int i = (Y + pbh - ty) / s;
            if (i < 0 || i >= ih) continue;
            for (int tx = (X + pbw) % s; tx < kw; tx += s) {
                const int j = (X + pbw - tx) / s;
                if (j < 0 || j >= iw) continue;
                const float* kp = kernel + (((long long)ty * kw + tx) * Cout + co) * Cin;
                const float* ip = inb + (long long)i * iw + j;
                float a = 0.f;
                for (int ci = 0; ci < Cin; ++ci) a = fmaf(__ldg(ip + (long long)ci * ih * iw), __ldg(kp + ci), a);
                acc += a;
            }
        }
        acc = apply_act(acc, act) * alpha;
        float* o = out + (long long)b * out_bstride + ((long long)co * oh + Y) * ow + X;
        *o = accumulate ? (*o + acc) : acc;
    }
}

// k == stride (every shipped config): non-overlapping transpose conv = one CoutxCin mat-vec per output
// pixel with a phase-selected matrix W[ty][tx].  A CTA owns one low-res row segment of `seg` pixels (the
// 32 x seg input tile sits in shared memory) and walks the s row phases ty; within a row phase every
// column phase tx is computed at once by a different thread group, so a pass yields one complete
// output row segment (seg*s pixels x Cout), staged phase-major in shared memory and then written
// (read-modify-written when accumulating) with fully coalesced rows.
// threads: tid -> (tx, cg, pg): 8 output channels x 4 low-res pixels per thread (32 accumulators), weights
// straight from global/L1 as float4 over ci (the same few KB for the whole CTA), inputs as float4 from smem.
__global__ void __launch_bounds__(288, 2) deconv_ks_kernel(const float* __restrict__ in, const float* __restrict__ kernel,
                                                           const float* __restrict__ bias, float* __restrict__ out, int Cin,
                                                           int Cout, int ih, int iw, int oh, int ow, int s, int pbh, int pbw,
                                                           int act, float alpha, int accumulate, long long out_bstride,
                                                           int seg, int phase_stride) {
    extern __shared__ __align__(16) float sm[];
    float* s_x = sm;                    // [32 ci][seg]
    float* s_o = sm + 32 * seg;         // [s tx][phase_stride >= 32 co * seg]; phase_stride % 32 == ceil(32/s): conflict-free gather
    const int b = blockIdx.z, i = blockIdx.y, j0 = blockIdx.x * seg;
    const int pgs = seg >> 2, nthr = blockDim.x;
    for (int e = threadIdx.x; e < 32 * seg; e += nthr) {
        const int ci = e / seg, px = e - ci * seg;
        s_x[e] = (ci < Cin && j0 + px < iw) ? __ldg(in + (((long long)b * Cin + ci) * ih + i) * iw + j0 + px) : 0.f;
    }
    const int tx = threadIdx.x / (4 * pgs);
    const int rem = threadIdx.x - tx * 4 * pgs;
    const int cg = rem / pgs, pg = rem - cg * pgs;
    const bool worker = tx < s;
    const int ncol = seg * s;           // output pixels of the row segment
    const int X0 = j0 * s - pbw;
    __syncthreads();
    for (int ty = 0; ty < s; ++ty) {
        const int Y = i * s + ty - pbh;
        if (Y < 0 || Y >= oh) continue;     // block-uniform
        float acc[8][4];
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[c][q] = 0.f;
        if (worker) {
            const float* wp = kernel + ((long long)(ty * s + tx) * Cout + cg * 8) * Cin;
            const float* xp = s_x + pg * 4;
#pragma unroll 1
            for (int ci = 0; ci < Cin; ci += 4) {
                float4 x4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) x4[k] = *reinterpret_cast<const float4*>(xp + (ci + k) * seg);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cg * 8 + c < Cout) w = __ldg(reinterpret_cast<const float4*>(wp + (long long)c * Cin + ci));
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
        }
        __syncthreads();                    // previous row phase's write-out finished with s_o
        if (worker) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(s_o + tx * phase_stride + (cg * 8 + c) * seg + pg * 4) =
                    make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
        }
        __syncthreads();
        // coalesced write-out: out[b, co, Y, X] (op)= alpha * act(v + bias); X = px*s + tx - pbw
        for (int e = threadIdx.x; e < Cout * ncol; e += nthr) {
            const int co = e / ncol, c = e - co * ncol;
            const int px = c / s, t = c - px * s;
            const int X = X0 + c;
            if (X < 0 || X >= ow || j0 + px >= iw) continue;
            float v = s_o[t * phase_stride + co * seg + px] + (bias ? __ldg(bias + co) : 0.f);
            v = apply_act(v, act) * alpha;
            float* o = out + (long long)b * out_bstride + ((long long)co * oh + Y) * ow + X;
            *o = accumulate ? (*o + v) : v;
        }
    }
}

// ------------------------------------------------------------------ tf.image.resize (separable gather)
__global__ void resize_kernel(const float* __restrict__ in, const int* __restrict__ iy,
                              const float* __restrict__ wy, const int* __restrict__ ix,
                              const float* __restrict__ wx, int taps, float* __restrict__ out, int C,
                              int ih, int iw, int oh, int ow, float alpha, int accumulate,
                              long long out_bstride, long long total) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int X = idx % ow;
        long long t = idx / ow;
        const int Y = t % oh; t /= oh;
        const int c = t % C;
        const int b = t / C;
        const float* src = in + ((long long)b * C + c) * ih * iw;
        float acc = 0.f;
        for (int a = 0; a < taps; ++a) {
            const float* row = src + (long long)__ldg(iy + Y * taps + a) * iw;
            float r = 0.f;   // TF interpolates along x first, then along y
            for (int q = 0; q < taps; ++q) r = fmaf(__ldg(row + __ldg(ix + X * taps + q)), __ldg(wx + X * taps + q), r);
            acc = fmaf(r, __ldg(wy + Y * taps + a), acc);
        }
        acc *= alpha;
        float* o = out + (long long)b * out_bstride + ((long long)c * oh + Y) * ow + X;
        *o = accumulate ? (*o + acc) : acc;
    }
}

// Small-source variant (the ds = 32/64/128 branches upsample 2x2..8x8 maps): the whole [C,ih,iw] source of
// one sample sits in shared memory and the read-modify-write of `out` is a coalesced stream.
__global__ void __launch_bounds__(256) resize_small_kernel(const float* __restrict__ in, const int* __restrict__ iy,
                                                           const float* __restrict__ wy, const int* __restrict__ ix,
                                                           const float* __restrict__ wx, int taps, float* __restrict__ out,
                                                           int C, int ih, int iw, int oh, int ow, int rows_per_cta,
                                                           float alpha, int accumulate, long long out_bstride) {
    extern __shared__ float ssrc[];
    const int b = blockIdx.y;
    const int Y0 = blockIdx.x * rows_per_cta;
    const int nsrc = C * ih * iw;
    for (int i = threadIdx.x; i < nsrc; i += blockDim.x) ssrc[i] = __ldg(in + (long long)b * nsrc + i);
    __syncthreads();
    const int rows = min(rows_per_cta, oh - Y0);
    for (int X = threadIdx.x; X < ow; X += blockDim.x) {
        int xi[4]; float xw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { xi[q] = q < taps ? __ldg(ix + X * taps + q) : 0; xw[q] = q < taps ? __ldg(wx + X * taps + q) : 0.f; }
        for (int r = 0; r < rows; ++r) {
            const int Y = Y0 + r;
            int yi[4]; float yw[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { yi[a] = a < taps ? __ldg(iy + Y * taps + a) * iw : 0; yw[a] = a < taps ? __ldg(wy + Y * taps + a) : 0.f; }
            for (int c = 0; c < C; ++c) {
                const float* src = ssrc + c * ih * iw;
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (a < taps) {
                        float rr = 0.f;   // TF interpolates along x first, then along y
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q < taps) rr = fmaf(src[yi[a] + xi[q]], xw[q], rr);
                        acc = fmaf(rr, yw[a], acc);
                    }
                }
                acc *= alpha;
                float* o = out + (long long)b * out_bstride + ((long long)c * oh + Y) * ow + X;
                *o = accumulate ? (*o + acc) : acc;
            }
        }
    }
}

// ------------------------------------------------------------------ spatial pyramid pooling
__global__ void spp_kernel(const float* __restrict__ in, const int* __restrict__ boxes,
                           float* __restrict__ out, int C, int H, int W, int nbins, int mode) {
    const int bin = blockIdx.x, b = blockIdx.y;
    const int y0 = boxes[bin * 4 + 0], y1 = boxes[bin * 4 + 1], x0 = boxes[bin * 4 + 2], x1 = boxes[bin * 4 + 3];
    const int bh = y1 - y0, bw = x1 - x0;
    const long long n = (long long)C * bh * bw;
    float acc = (mode == PCNN_POOL_MAX) ? -INFINITY : 0.f;
    const float* src = in + (long long)b * C * H * W;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const int x = i % bw;
        long long t = i / bw;
        const int y = t % bh;
        const int c = t / bh;
        const float v = __ldg(src + ((long long)c * H + (y0 + y)) * W + x0 + x);
        acc = (mode == PCNN_POOL_MAX) ? fmaxf(acc, v) : acc + v;
    }
    __shared__ float red[32];
    acc = (mode == PCNN_POOL_MAX) ? warp_max(acc) : warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        float v = (threadIdx.x < nw) ? red[threadIdx.x] : ((mode == PCNN_POOL_MAX) ? -INFINITY : 0.f);
        v = (mode == PCNN_POOL_MAX) ? warp_max(v) : warp_sum(v);
        if (threadIdx.x == 0) out[(long long)b * nbins + bin] = (mode == PCNN_POOL_MAX) ? v : v / (float)n;   // n==0 -> NaN like reduce_mean
    }
}

// ------------------------------------------------------------------ Dense
__global__ void dense_kernel(const float* __restrict__ x, const float* __restrict__ k,
                             const float* __restrict__ bias, float* __restrict__ y, int B, int nin,
                             int nout, int act) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= nout) return;
    float acc = 0.f;
    const float* xb = x + (long long)b * nin;
    for (int i = 0; i < nin; ++i) acc = fmaf(__ldg(xb + i), __ldg(k + (long long)i * nout + j), acc);
    if (bias) acc += __ldg(bias + j);
    y[(long long)b * nout + j] = apply_act(acc, act);
}

// ------------------------------------------------------------------ max |x| per sample
__global__ void maxabs_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
    const int b = blockIdx.y;
    const float* src = x + (long long)b * n;
    float m = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(src + i)));
    __shared__ float red[32];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        v = warp_max(v);
        // non-negative floats order like their bit patterns; NaN inputs are not expected here
        if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(out + b), __float_as_int(v));
    }
}

__global__ void scale_inv_kernel(const float* __restrict__ x, const float* __restrict__ maxabs,
                                 float* __restrict__ y, long long n) {
    const int b = blockIdx.y;
    const float f = 1.0f / __ldg(maxabs + b);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[(long long)b * n + i] = __ldg(x + (long long)b * n + i) * f;
}

// ------------------------------------------------------------------ model glue
__global__ void hpnn_input_kernel(const float* __restrict__ rhs, const float* __restrict__ posx,
                                  const float* __restrict__ posy, float* __restrict__ out, int H, int W,
                                  long long total) {
    const long long plane = (long long)H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / plane, r = idx - b * plane;
        const int i = r / W, j = r - (long long)i * W;
        float* o = out + b * 3 * plane + r;
        o[0] = __ldg(rhs + idx);
        o[plane] = __ldg(posx + i);
        o[2 * plane] = __ldg(posy + j);
    }
}

__global__ void dbcnn_input_kernel(const float* __restrict__ bc, float posx0, const float* __restrict__ posy,
                                   float* __restrict__ out, int n, long long total) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / n;
        const int j = idx - b * n;
        float* o = out + b * 3 * n + j;
        o[0] = __ldg(bc + idx);
        o[n] = posx0;
        o[2 * n] = __ldg(posy + j);
    }
}

__global__ void dbcnn_expand_kernel(const float* __restrict__ h, const float* __restrict__ S,
                                    const float* __restrict__ mw, const float* __restrict__ posx,
                                    const float* __restrict__ posy, float* __restrict__ out, int M,
                                    int xres, int n, long long total) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int y = idx % n;
        long long t = idx / n;
        const int x = t % xres; t /= xres;
        const int m = t % (M + 2);
        const long long b = t / (M + 2);
        float v;
        if (m < M) v = __ldg(h + (b * M + m) * n + y) * __ldg(S + (long long)m * xres + x) * __ldg(mw + b * M + m);
        else if (m == M) v = __ldg(posx + x);
        else v = __ldg(posy + y);
        out[idx] = v;
    }
}

__global__ void dbcnn_finalize_kernel(const float* __restrict__ raw, const float* __restrict__ maxabs,
                                      const float* __restrict__ bc, float* __restrict__ out, int xres,
                                      int n, long long total) {
    const long long plane = (long long)xres * n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / plane, r = idx - b * plane;
        out[idx] = (r < n) ? __ldg(bc + b * n + r) : __ldg(raw + idx) * (1.0f / __ldg(maxabs + b));
    }
}

__global__ void hpnn_finalize_kernel(const float* __restrict__ y, const float* __restrict__ s,
                                     float* __restrict__ out, int H, int W, int bc_type,
                                     long long y_bstride, long long total) {
    const long long plane = (long long)H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / plane, r = idx - b * plane;
        const int i = r / W, j = r - (long long)i * W;
        const float f = s ? (1.0f + __ldg(s + b)) : 1.0f;
        const bool ring = (i == 0) | (i == H - 1) | (j == 0) | (j == W - 1);
        float v;
        if (!ring) v = f * __ldg(y + b * y_bstride + r);
        else if (bc_type == PCNN_BC_DIRICHLET) v = 0.f;
        else {
            const int ii = min(max(i, 1), H - 2), jj = min(max(j, 1), W - 2);
            v = f * __ldg(y + b * y_bstride + (long long)ii * W + jj);
        }
        out[idx] = v;
    }
}

// dense_inp of the two MLPs: [dx, dx*(H-1), dx*(W-1)] (HPNN, Homogeneous_Poisson_NN_Legacy.py:193,202)
// or [dx, Lx/Lmax, Ly/Lmax, spp...] (DBCNN, Dirichlet_BC_NN_Legacy.py:129-130,142)
__global__ void dense_input_kernel(const float* __restrict__ dx, const float* __restrict__ extra,
                                   float* __restrict__ out, int B, int n0, int n1, int nextra, int normalize) {
    const int width = 3 + nextra;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B * width; idx += gridDim.x * blockDim.x) {
        const int b = idx / width, k = idx - b * width;
        const float d = __ldg(dx + b);
        const float l0 = d * (float)(n0 - 1), l1 = d * (float)(n1 - 1);
        const float lm = fmaxf(l0, l1);
        float v;
        if (k == 0) v = d;
        else if (k == 1) v = normalize ? l0 / lm : l0;
        else if (k == 2) v = normalize ? l1 / lm : l1;
        else v = __ldg(extra + (long long)b * nextra + (k - 3));
        out[idx] = v;
    }
}

__global__ void merge_kernel(const float* __restrict__ hp, const float* __restrict__ L,
                             const float* __restrict__ T, const float* __restrict__ R,
                             const float* __restrict__ Bt, const float* __restrict__ dx,
                             const float* __restrict__ mrhs, const float* __restrict__ ml,
                             const float* __restrict__ mt, const float* __restrict__ mr,
                             const float* __restrict__ mb, float* __restrict__ out, int nx, int ny,
                             long long total) {
    const long long plane = (long long)nx * ny;
    const float nmax = (float)(max(nx, ny) - 1);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / plane, r = idx - b * plane;
        const int i = r / ny, j = r - (long long)i * ny;
        const float lmax = __ldg(dx + b) * nmax;
        const float fh = lmax * lmax * __ldg(mrhs + b);   // max_domain_size^2 / rhs_scaling_factor
        const float vl = __ldg(L + b * plane + (long long)i * ny + j) * __ldg(ml + b);
        const float vr = __ldg(R + b * plane + (long long)(nx - 1 - i) * ny + j) * __ldg(mr + b);
        const float vt = __ldg(T + b * plane + (long long)(ny - 1 - j) * nx + i) * __ldg(mt + b);
        const float vb = __ldg(Bt + b * plane + (long long)j * nx + i) * __ldg(mb + b);
        out[idx] = vl + vr + vt + vb + __ldg(hp + idx) * fh;   // reference order: left+right+top+bottom+hpnn
    }
}

static inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
    long long g = (total + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace pcnn

using namespace pcnn;

extern "C" int pcnn_version(void) { return PCNN_VERSION; }
extern "C" const char* pcnn_last_error(void) { return pcnn::g_err; }
extern "C" long long pcnn_launch_count(void) { return pcnn::g_launches.load(); }

extern "C" int pcnn_avgpool_same_f32(const float* in, float* out, int B, int C, int H, int W, int s,
                                     int64_t in_bstride, void* stream) {
    PCNN_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0 && s > 0, "avgpool_same_f32: bad argument");
    const int oh = ceil_div(H, s), ow = ceil_div(W, s);
    const int pt = (oh * s - H) / 2, pl = (ow * s - W) / 2;
    PCNN_CHECK_ARG(oh <= 65535 && (long long)B * C <= 65535, "avgpool_same_f32: grid too large");
    if (s <= 4) {
        avgpool_thread_kernel<<<dim3(ceil_div(ow, 128), oh, B * C), 128, 0, (cudaStream_t)stream>>>(in, out, C, H, W, oh, ow, s, pt, pl, in_bstride);
    } else {
        PCNN_CHECK_ARG(W <= 12288, "avgpool_same_f32: rows wider than 12288 are not supported for s > 4");
        const int rows_per_cta = std::max(1, 32 / s);          // >= 32 input rows per CTA
        avgpool_rows_kernel<<<dim3(ceil_div(oh, rows_per_cta), B * C), 256, (size_t)W * sizeof(float), (cudaStream_t)stream>>>(in, out, C, H, W, oh, ow, s, pt, pl, in_bstride, rows_per_cta);
    }
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_deconv_same_f32(const float* in, const float* kernel, const float* bias, float* out,
                                    int B, int Cin, int Cout, int ih, int iw, int oh, int ow, int kh,
                                    int kw, int stride, int act, float alpha, int accumulate,
                                    int64_t out_bstride, void* stream) {
    PCNN_CHECK_ARG(in && kernel && out && B > 0 && Cin > 0 && Cout > 0 && stride > 0, "deconv_same_f32: bad argument");
    PCNN_CHECK_ARG(ceil_div(oh, stride) == ih && ceil_div(ow, stride) == iw,
                   "deconv_same_f32: output_shape (%d,%d) inconsistent with input (%d,%d) at stride %d (TF raises)", oh, ow, ih, iw, stride);
    const int pbh = max((ih - 1) * stride + kh - oh, 0) / 2;
    const int pbw = max((iw - 1) * stride + kw - ow, 0) / 2;
    if (kh == stride && kw == stride && Cin <= 32 && (Cin % 4) == 0 && Cout <= 32 && stride <= 16 && B <= 65535 && ih <= 65535 &&
        (reinterpret_cast<uintptr_t>(kernel) % 16) == 0) {
        // low-res pixels per CTA: seg*stride ~ 256 output pixels, seg a multiple of 4; all `stride` column phases run at once
        int seg = ((256 / stride + 3) / 4) * 4;
        seg = std::min(seg, ((iw + 3) / 4) * 4);
        const int threads = ((seg * stride + 31) / 32) * 32;        // 4 channel groups x seg/4 pixel groups x stride phases
        int phase_stride = 32 * seg;                                // == 0 mod 32
        phase_stride += (32 + stride - 1) / stride;                 // phases land ceil(32/s) banks apart
        phase_stride = (phase_stride + 3) & ~3;                     // float4 staging stores stay aligned
        const size_t smem = ((size_t)32 * seg + (size_t)stride * phase_stride) * sizeof(float);
        if (threads <= 288 && smem <= 100 * 1024) {
            if (smem > 48 * 1024)
                PCNN_CHECK_CUDA(cudaFuncSetAttribute(deconv_ks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid(ceil_div(iw, seg), ih, B);
            deconv_ks_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(in, kernel, bias, out, Cin, Cout, ih, iw, oh, ow, stride, pbh, pbw, act, alpha, accumulate, out_bstride, seg, phase_stride);
            PCNN_CHECK_LAUNCH();
            return PCNN_OK;
        }
    }
    const long long total = (long long)B * Cout * oh * ow;
    deconv_same_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(in, kernel, bias, out, Cin, Cout, ih, iw, oh, ow, kh, kw, stride, pbh, pbw, act, alpha, accumulate, out_bstride, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_resize_f32(const float* in, const int32_t* iy, const float* wy, const int32_t* ix,
                               const float* wx, int taps, float* out, int B, int C, int ih, int iw,
                               int oh, int ow, float alpha, int accumulate, int64_t out_bstride,
                               void* stream) {
    PCNN_CHECK_ARG(in && iy && wy && ix && wx && out && taps >= 1 && taps <= 4, "resize_f32: bad argument");
    if ((size_t)C * ih * iw * sizeof(float) <= 32 * 1024 && B <= 65535) {
        const int rows_per_cta = 4;
        resize_small_kernel<<<dim3(ceil_div(oh, rows_per_cta), B), 256, (size_t)C * ih * iw * sizeof(float), (cudaStream_t)stream>>>(
            in, iy, wy, ix, wx, taps, out, C, ih, iw, oh, ow, rows_per_cta, alpha, accumulate, out_bstride);
        PCNN_CHECK_LAUNCH();
        return PCNN_OK;
    }
    const long long total = (long long)B * C * oh * ow;
    resize_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(in, iy, wy, ix, wx, taps, out, C, ih, iw, oh, ow, alpha, accumulate, out_bstride, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_spp_f32(const float* in, const int32_t* boxes, float* out, int B, int C, int H, int W,
                            int nbins, int mode, void* stream) {
    PCNN_CHECK_ARG(in && boxes && out && B > 0 && nbins > 0 && B <= 65535, "spp_f32: bad argument");
    spp_kernel<<<dim3(nbins, B), 128, 0, (cudaStream_t)stream>>>(in, boxes, out, C, H, W, nbins, mode);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_dense_f32(const float* x, const float* kernel, const float* bias, float* y, int B,
                              int nin, int nout, int act, void* stream) {
    PCNN_CHECK_ARG(x && kernel && y && B > 0 && nin > 0 && nout > 0 && B <= 65535, "dense_f32: bad argument");
    dense_kernel<<<dim3(ceil_div(nout, 128), B), 128, 0, (cudaStream_t)stream>>>(x, kernel, bias, y, B, nin, nout, act);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_maxabs_f32(const float* x, float* maxabs, int B, int64_t n, void* stream) {
    PCNN_CHECK_ARG(x && maxabs && B > 0 && n > 0 && B <= 65535, "maxabs_f32: bad argument");
    PCNN_CHECK_CUDA(cudaMemsetAsync(maxabs, 0, sizeof(float) * B, (cudaStream_t)stream));
    int gx = (int)std::min<long long>((n + 2047) / 2048, 64);
    maxabs_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(x, maxabs, n);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_scale_inv_f32(const float* x, const float* maxabs, float* y, int B, int64_t n, void* stream) {
    PCNN_CHECK_ARG(x && maxabs && y && B > 0 && n > 0 && B <= 65535, "scale_inv_f32: bad argument");
    int gx = (int)std::min<long long>((n + 1023) / 1024, 64);
    scale_inv_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(x, maxabs, y, n);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_hpnn_input_f32(const float* rhs, const float* posx, const float* posy, float* out,
                                   int B, int H, int W, void* stream) {
    PCNN_CHECK_ARG(rhs && posx && posy && out && B > 0, "hpnn_input_f32: bad argument");
    const long long total = (long long)B * H * W;
    hpnn_input_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(rhs, posx, posy, out, H, W, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_dbcnn_input_f32(const float* bc, float posx0, const float* posy, float* out, int B,
                                    int n, void* stream) {
    PCNN_CHECK_ARG(bc && posy && out && B > 0 && n > 0, "dbcnn_input_f32: bad argument");
    const long long total = (long long)B * n;
    dbcnn_input_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(bc, posx0, posy, out, n, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_dbcnn_expand_f32(const float* h, const float* sinh_basis, const float* modew,
                                     const float* posx, const float* posy, float* out, int B, int M,
                                     int xres, int n, void* stream) {
    PCNN_CHECK_ARG(h && sinh_basis && modew && posx && posy && out && B > 0 && M > 0, "dbcnn_expand_f32: bad argument");
    const long long total = (long long)B * (M + 2) * xres * n;
    dbcnn_expand_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(h, sinh_basis, modew, posx, posy, out, M, xres, n, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_dbcnn_finalize_f32(const float* raw, const float* maxabs, const float* bc, float* out,
                                       int B, int xres, int n, void* stream) {
    PCNN_CHECK_ARG(raw && maxabs && bc && out && B > 0, "dbcnn_finalize_f32: bad argument");
    const long long total = (long long)B * xres * n;
    dbcnn_finalize_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(raw, maxabs, bc, out, xres, n, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_dense_input_f32(const float* dx, const float* extra, float* out, int B, int n0, int n1,
                                    int nextra, int normalize, void* stream) {
    PCNN_CHECK_ARG(dx && out && B > 0 && nextra >= 0 && (nextra == 0 || extra), "dense_input_f32: bad argument");
    dense_input_kernel<<<grid_for((long long)B * (3 + nextra)), 256, 0, (cudaStream_t)stream>>>(dx, extra, out, B, n0, n1, nextra, normalize);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_hpnn_finalize_f32(const float* y, const float* s, float* out, int B, int H, int W,
                                      int bc_type, int64_t y_bstride, void* stream) {
    PCNN_CHECK_ARG(y && out && B > 0 && H >= 3 && W >= 3, "hpnn_finalize_f32: bad argument");
    PCNN_CHECK_ARG(bc_type == PCNN_BC_DIRICHLET || bc_type == PCNN_BC_NEUMANN, "bc_type can only be neumann or dirichlet.");
    const long long total = (long long)B * H * W;
    hpnn_finalize_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(y, s, out, H, W, bc_type, y_bstride, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_merge_f32(const float* hp, const float* L, const float* T, const float* R,
                              const float* Bt, const float* dx, const float* mrhs, const float* ml,
                              const float* mt, const float* mr, const float* mb, float* out, int B,
                              int nx, int ny, void* stream) {
    PCNN_CHECK_ARG(hp && L && T && R && Bt && dx && mrhs && ml && mt && mr && mb && out && B > 0, "merge_f32: bad argument");
    const long long total = (long long)B * nx * ny;
    merge_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(hp, L, T, R, Bt, dx, mrhs, ml, mt, mr, mb, out, nx, ny, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
