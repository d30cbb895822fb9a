// Accuracy-check kernels: finite-difference Laplacian residual (+ per-sample squared norm via
// warp-shuffle reductions), Jacobi sweep, and the DST-I direct Poisson solve that supplies ground
// truth.  The residual kernel is HBM-bound at 8 B/grid point (read u, read f).
#include "pcnn_common.cuh"

namespace pcnn {

// ------------------------------------------------------------------ Laplacian residual
// One warp owns a strip of RS interior rows x 128 columns: every lane marches down the strip with 4 consecutive columns
// (one 16-byte load per row of u and of f) and a rolling register window of 2*HALO+1 rows; the left/right neighbours come
// from the adjacent lanes by shuffle (the two edge lanes load them).  No index division per element; DRAM traffic is
// (RS + 2*HALO)/RS x 4 B for u plus 4 B for f per grid point (RS = 32: 8.25 B / 8.5 B against the 8 B algorithmic figure).
// Squared residuals are accumulated in double per lane, reduced by warp shuffles, one double atomic per CTA.
template <bool VEC>
__device__ __forceinline__ float4 load_row4(const float* __restrict__ row, int j0, int W) {
    if (VEC) {
        if (j0 + 3 < W) return __ldg(reinterpret_cast<const float4*>(row + j0));
    }
    float4 v;
    v.x = (j0 + 0 < W) ? __ldg(row + j0 + 0) : 0.f;
    v.y = (j0 + 1 < W) ? __ldg(row + j0 + 1) : 0.f;
    v.z = (j0 + 2 < W) ? __ldg(row + j0 + 2) : 0.f;
    v.w = (j0 + 3 < W) ? __ldg(row + j0 + 3) : 0.f;
    return v;
}

constexpr int kResidualStripRows = 32;

template <int HALO, bool VEC>
__global__ void __launch_bounds__(256) laplacian_residual_kernel(
    const float* __restrict__ rhs, const float* __restrict__ sol, const float* __restrict__ gs,
    const float* __restrict__ rhs_maxabs, double* __restrict__ sq_sum, int H, int W, int nchunk, int nitems) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float qx = 1.0f / (gs[b * 2 + 0] * gs[b * 2 + 0]);
    const float qy = 1.0f / (gs[b * 2 + 1] * gs[b * 2 + 1]);
    // central second-derivative coefficients (dataset/utils/get_fd_coefficients.py)
    float c[5];
    if constexpr (HALO == 1) { c[0] = 1.f; c[1] = -2.f; c[2] = 1.f; c[3] = 0.f; c[4] = 0.f; }
    else { c[0] = -1.f / 12.f; c[1] = 4.f / 3.f; c[2] = -2.5f; c[3] = 4.f / 3.f; c[4] = -1.f / 12.f; }
    const float cc = c[HALO] * qx + c[HALO] * qy;             // kernel centre = sum_d stencil_d*q_d
    const float* u = sol + (long long)b * H * W;
    const float* f = rhs + (long long)b * H * W;
    double acc = 0.0;
    const int item = blockIdx.x * 8 + warp;
    if (item < nitems) {
        const int chunk = item % nchunk, strip = item / nchunk;
        const int j0 = chunk * 128 + lane * 4;
        const int i0 = HALO + strip * kResidualStripRows;
        const int i1 = min(i0 + kResidualStripRows, H - HALO);
        float4 win[2 * HALO + 1];
#pragma unroll
        for (int t = 0; t < 2 * HALO; ++t) win[t + 1] = load_row4<VEC>(u + (long long)(i0 - HALO + t) * W, j0, W);
        for (int i = i0; i < i1; ++i) {
#pragma unroll
            for (int t = 0; t < 2 * HALO; ++t) win[t] = win[t + 1];
            win[2 * HALO] = load_row4<VEC>(u + (long long)(i + HALO) * W, j0, W);
            const float4 fv = load_row4<VEC>(f + (long long)i * W, j0, W);
            const float4 mid = win[HALO];
            // columns j0-HALO .. j0+3+HALO of the centre row
            float row[4 + 2 * HALO];
            row[HALO + 0] = mid.x; row[HALO + 1] = mid.y; row[HALO + 2] = mid.z; row[HALO + 3] = mid.w;
            const float* ur = u + (long long)i * W;
            {
                float l1 = __shfl_up_sync(0xffffffffu, mid.w, 1), r1 = __shfl_down_sync(0xffffffffu, mid.x, 1);
                if (lane == 0) l1 = (j0 >= 1) ? __ldg(ur + j0 - 1) : 0.f;
                if (lane == 31) r1 = (j0 + 4 < W) ? __ldg(ur + j0 + 4) : 0.f;
                row[HALO - 1] = l1; row[HALO + 4] = r1;
            }
            if constexpr (HALO == 2) {
                float l2 = __shfl_up_sync(0xffffffffu, mid.z, 1), r2 = __shfl_down_sync(0xffffffffu, mid.y, 1);
                if (lane == 0) l2 = (j0 >= 2) ? __ldg(ur + j0 - 2) : 0.f;
                if (lane == 31) r2 = (j0 + 5 < W) ? __ldg(ur + j0 + 5) : 0.f;
                row[0] = l2; row[HALO + 5] = r2;
            }
            const float fe[4] = {fv.x, fv.y, fv.z, fv.w};
            float col[2 * HALO + 1][4];
#pragma unroll
            for (int t = 0; t <= 2 * HALO; ++t) { col[t][0] = win[t].x; col[t][1] = win[t].y; col[t][2] = win[t].z; col[t][3] = win[t].w; }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + e;
                float lap = cc * row[HALO + e];
#pragma unroll
                for (int t = 1; t <= HALO; ++t) {
                    lap = fmaf(c[HALO - t] * qx, col[HALO - t][e], lap);
                    lap = fmaf(c[HALO + t] * qx, col[HALO + t][e], lap);
                    lap = fmaf(c[HALO - t] * qy, row[HALO + e - t], lap);
                    lap = fmaf(c[HALO + t] * qy, row[HALO + e + t], lap);
                }
                const float d = fe[e] - lap;
                if (j >= HALO && j < W - HALO) acc += (double)d * (double)d;
            }
        }
    }
    __shared__ double red[8];
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            if (rhs_maxabs) { const double m = (double)rhs_maxabs[b]; v /= (m * m); }
            atomicAdd(sq_sum + b, v);
        }
    }
}

__global__ void jacobi_sweep_kernel(const float* __restrict__ cur, const float* __restrict__ rhs,
                                    const float* __restrict__ gs, float* __restrict__ next, int H, int W,
                                    long long total) {
    const long long plane = (long long)H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / plane, r = idx - b * plane;
        const int i = r / W, j = r - (long long)i * W;
        float v = __ldg(cur + idx);
        if (i > 0 && i < H - 1 && j > 0 && j < W - 1) {
            const float px = 1.0f / (gs[b * 2] * gs[b * 2]), py = 1.0f / (gs[b * 2 + 1] * gs[b * 2 + 1]);
            const float dinv = 1.0f / (-2.0f * px - 2.0f * py);
            const float cr = px * (__ldg(cur + idx - W) + __ldg(cur + idx + W)) + py * (__ldg(cur + idx - 1) + __ldg(cur + idx + 1));
            v = dinv * (__ldg(rhs + idx) - cr);
        }
        next[idx] = v;
    }
}

// float4 variant (W % 4 == 0, 16-byte aligned maps): one thread per 4 consecutive points of a row, one 32-bit division per 4
// points; the arithmetic per point is the scalar kernel's.
__global__ void __launch_bounds__(256) jacobi_sweep_vec4_kernel(const float* __restrict__ cur, const float* __restrict__ rhs,
                                                               const float* __restrict__ gs, float* __restrict__ next, int H, int W4) {
    const int b = blockIdx.y;
    const unsigned n4 = (unsigned)H * (unsigned)W4;
    const long long base = (long long)b * n4 * 4;
    const float* cb = cur + base;
    const float4* c4 = reinterpret_cast<const float4*>(cb);
    const float4* f4 = reinterpret_cast<const float4*>(rhs + base);
    float4* o4 = reinterpret_cast<float4*>(next + base);
    const float px = 1.0f / (gs[b * 2] * gs[b * 2]), py = 1.0f / (gs[b * 2 + 1] * gs[b * 2 + 1]);
    const float dinv = 1.0f / (-2.0f * px - 2.0f * py);
    for (unsigned v = blockIdx.x * 256u + threadIdx.x; v < n4; v += gridDim.x * 256u) {
        const unsigned i = v / (unsigned)W4, j4 = v - i * (unsigned)W4;
        const float4 c = __ldg(c4 + v);
        float4 o = c;
        if (i > 0 && i < (unsigned)H - 1) {
            const float4 up = __ldg(c4 + v - W4), dn = __ldg(c4 + v + W4), f = __ldg(f4 + v);
            if (j4 > 0) o.x = dinv * (f.x - (px * (up.x + dn.x) + py * (__ldg(cb + (long long)v * 4 - 1) + c.y)));
            o.y = dinv * (f.y - (px * (up.y + dn.y) + py * (c.x + c.z)));
            o.z = dinv * (f.z - (px * (up.z + dn.z) + py * (c.y + c.w)));
            if (j4 < (unsigned)W4 - 1) o.w = dinv * (f.w - (px * (up.w + dn.w) + py * (c.z + __ldg(cb + (long long)v * 4 + 4))));
        }
        o4[v] = o;
    }
}

// ------------------------------------------------------------------ DST-I direct solve (double)
__global__ void dst_sine_kernel(double* s, int m) {
    const long long total = (long long)m * m;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long j = idx / m + 1, k = idx % m + 1;
        const long long r = (j * k) % (2LL * (m + 1));          // exact argument reduction
        s[idx] = sinpi((double)r / (double)(m + 1));
    }
}

// b = -dx^2 f + adjacent Dirichlet values, interior only  (dataset/solvers/cholesky.py:45-119)
__global__ void dst_build_rhs_kernel(const float* __restrict__ rhs, const float* __restrict__ left,
                                     const float* __restrict__ top, const float* __restrict__ right,
                                     const float* __restrict__ bottom, const float* __restrict__ dx,
                                     double* __restrict__ bvec, int nx, int ny, long long total) {
    const int mx = nx - 2, my = ny - 2;
    const long long plane = (long long)mx * my;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / plane, r = idx - b * plane;
        const int i = r / my + 1, j = r % my + 1;
        const double h = (double)dx[b];
        double v = -h * h * (double)rhs[b * nx * ny + (long long)i * ny + j];
        if (j == 1) v += (double)bottom[b * nx + i];
        if (j == ny - 2) v += (double)top[b * nx + i];
        if (i == 1) v += (double)left[b * ny + j];
        if (i == nx - 2) v += (double)right[b * ny + j];
        bvec[idx] = v;
    }
}

// C[b] = A[b] * Bm[b]  (strideA or strideB may be 0 for a shared operand); row-major; 32x32 tiles.
__global__ void __launch_bounds__(256) dgemm_batched_kernel(const double* __restrict__ A, long long sA,
                                                            const double* __restrict__ Bm, long long sB,
                                                            double* __restrict__ C, long long sC, int M,
                                                            int N, int K) {
    __shared__ double As[32][17];
    __shared__ double Bs[16][33];
    const int b = blockIdx.z;
    const double* a = A + b * sA;
    const double* bm = Bm + b * sB;
    double* c = C + b * sC;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
    double acc[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int e = threadIdx.x; e < 32 * 16; e += 256) {
            const int r = e / 16, kk = e % 16;
            As[r][kk] = (row0 + r < M && k0 + kk < K) ? a[(long long)(row0 + r) * K + k0 + kk] : 0.0;
            const int kr = e / 32, cc = e % 32;
            Bs[kr][cc] = (k0 + kr < K && col0 + cc < N) ? bm[(long long)(k0 + kr) * N + col0 + cc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const double a0 = As[ty][kk], a1 = As[ty + 16][kk];
            const double b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
            acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = row0 + ty + 16 * i, cc = col0 + tx + 16 * j;
            if (r < M && cc < N) c[(long long)r * N + cc] = acc[i][j];
        }
}

__global__ void dst_eigen_scale_kernel(double* __restrict__ v, int mx, int my, long long total) {
    const long long plane = (long long)mx * my;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx % plane;
        const int i = r / my + 1, j = r % my + 1;
        const double lam = (2.0 - 2.0 * cospi((double)i / (mx + 1))) + (2.0 - 2.0 * cospi((double)j / (my + 1)));
        v[idx] /= lam;
    }
}

__global__ void dst_write_solution_kernel(const double* __restrict__ u, const float* __restrict__ left,
                                          const float* __restrict__ top, const float* __restrict__ right,
                                          const float* __restrict__ bottom, float* __restrict__ out, int nx,
                                          int ny, long long total) {
    const int mx = nx - 2, my = ny - 2;
    const long long plane = (long long)nx * ny;
    const double norm = (2.0 / (mx + 1)) * (2.0 / (my + 1));
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / plane, r = idx - b * plane;
        const int i = r / ny, j = r % ny;
        float v;
        // write order of multigrid.py:145-148: top, bottom, then left, right (corners from left/right)
        if (i == 0) v = left[b * ny + j];
        else if (i == nx - 1) v = right[b * ny + j];
        else if (j == 0) v = bottom[b * nx + i];
        else if (j == ny - 1) v = top[b * nx + i];
        else v = (float)(u[b * mx * my + (long long)(i - 1) * my + (j - 1)] * norm);
        out[idx] = v;
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

extern "C" int pcnn_laplacian_residual_f32(const float* rhs, const float* sol, const float* grid_spacings,
                                           const float* rhs_maxabs, double* sq_sum, int B, int H, int W,
                                           int stencil, void* stream) {
    PCNN_CHECK_ARG(rhs && sol && grid_spacings && sq_sum && B > 0 && B <= 65535, "laplacian_residual_f32: bad argument");
    PCNN_CHECK_ARG(stencil == 3 || stencil == 5, "laplacian_residual_f32: stencil size %d not supported (3 or 5)", stencil);
    PCNN_CHECK_ARG(H > stencil - 1 && W > stencil - 1, "laplacian_residual_f32: grid smaller than the stencil");
    PCNN_CHECK_CUDA(cudaMemsetAsync(sq_sum, 0, sizeof(double) * B, (cudaStream_t)stream));
    const int halo = stencil / 2;
    const int nchunk = ceil_div(W, 128), nstrip = ceil_div(H - 2 * halo, kResidualStripRows), nitems = nchunk * nstrip;
    const dim3 grid(ceil_div(nitems, 8), B);
    const bool vec = (W % 4 == 0) && (((uintptr_t)rhs | (uintptr_t)sol) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (stencil == 3) {
        if (vec) laplacian_residual_kernel<1, true><<<grid, 256, 0, st>>>(rhs, sol, grid_spacings, rhs_maxabs, sq_sum, H, W, nchunk, nitems);
        else laplacian_residual_kernel<1, false><<<grid, 256, 0, st>>>(rhs, sol, grid_spacings, rhs_maxabs, sq_sum, H, W, nchunk, nitems);
    } else {
        if (vec) laplacian_residual_kernel<2, true><<<grid, 256, 0, st>>>(rhs, sol, grid_spacings, rhs_maxabs, sq_sum, H, W, nchunk, nitems);
        else laplacian_residual_kernel<2, false><<<grid, 256, 0, st>>>(rhs, sol, grid_spacings, rhs_maxabs, sq_sum, H, W, nchunk, nitems);
    }
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_jacobi_sweep_f32(const float* cur, const float* rhs, const float* grid_spacings,
                                     float* next, int B, int H, int W, void* stream) {
    PCNN_CHECK_ARG(cur && rhs && grid_spacings && next && cur != next && B > 0 && H >= 3 && W >= 3, "jacobi_sweep_f32: bad argument");
    const long long total = (long long)B * H * W;
    const long long n4 = (long long)H * (W / 4);
    if (W % 4 == 0 && B <= 65535 && n4 < (1ll << 31) && (((uintptr_t)cur | (uintptr_t)rhs | (uintptr_t)next) & 15) == 0) {
        const int gx = (int)std::min<long long>((n4 + 255) / 256, 148 * 16);
        jacobi_sweep_vec4_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(cur, rhs, grid_spacings, next, H, W / 4);
    } else {
        jacobi_sweep_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(cur, rhs, grid_spacings, next, H, W, total);
    }
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" size_t pcnn_dst_workspace_bytes(int B, int nx, int ny) {
    if (B <= 0 || nx < 3 || ny < 3) return 0;
    return (size_t)2 * B * (nx - 2) * (ny - 2) * sizeof(double);
}

extern "C" int pcnn_dst_sine_matrix(double* s, int m, void* stream) {
    PCNN_CHECK_ARG(s && m > 0, "dst_sine_matrix: bad argument");
    dst_sine_kernel<<<grid_for((long long)m * m), 256, 0, (cudaStream_t)stream>>>(s, m);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_dst_solve(const float* rhs, const float* left, const float* top, const float* right,
                              const float* bottom, const float* dx, const double* sx, const double* sy,
                              double* work, float* out, int B, int nx, int ny, void* stream) {
    PCNN_CHECK_ARG(rhs && left && top && right && bottom && dx && sx && sy && work && out, "dst_solve: null pointer");
    PCNN_CHECK_ARG(B > 0 && B <= 65535 && nx >= 3 && ny >= 3, "dst_solve: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int mx = nx - 2, my = ny - 2;
    const long long plane = (long long)mx * my, total = plane * B;
    double* w0 = work;
    double* w1 = work + total;
    dst_build_rhs_kernel<<<grid_for(total), 256, 0, st>>>(rhs, left, top, right, bottom, dx, w0, nx, ny, total);
    PCNN_CHECK_LAUNCH();
    dim3 grid(ceil_div(my, 32), ceil_div(mx, 32), B);
    dgemm_batched_kernel<<<grid, 256, 0, st>>>(sx, 0, w0, plane, w1, plane, mx, my, mx);   // Sx * b
    PCNN_CHECK_LAUNCH();
    dgemm_batched_kernel<<<grid, 256, 0, st>>>(w1, plane, sy, 0, w0, plane, mx, my, my);   // (.) * Sy
    PCNN_CHECK_LAUNCH();
    dst_eigen_scale_kernel<<<grid_for(total), 256, 0, st>>>(w0, mx, my, total);
    PCNN_CHECK_LAUNCH();
    dgemm_batched_kernel<<<grid, 256, 0, st>>>(sx, 0, w0, plane, w1, plane, mx, my, mx);
    PCNN_CHECK_LAUNCH();
    dgemm_batched_kernel<<<grid, 256, 0, st>>>(w1, plane, sy, 0, w0, plane, mx, my, my);
    PCNN_CHECK_LAUNCH();
    const long long tout = (long long)B * nx * ny;
    dst_write_solution_kernel<<<grid_for(tout), 256, 0, st>>>(w0, left, top, right, bottom, out, nx, ny, tout);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
