// The step AFTER the hot path: the pressure-Poisson solve of the reference's Navier-Stokes projection
// solver, which takes the Neumann HPNN's prediction as the initial guess of a Krylov iteration
// (Navier_Stokes_2D/solvers.py:29-33 builds the model, :153-186 the operator, :204-334 the solve).
//
// Operator (Poisson_pressure_matrix, solvers.py:164-173): A = kron(I, T) + kron(T, I), T = tridiag(-1, [1,2,...,2,1], -1)/dh^2,
// i.e. minus the cell-centred 5-point Laplacian with homogeneous Neumann boundaries (the diagonal counts the neighbours
// that exist).  A is symmetric positive semi-definite with the constants as null space; the reference pins the constant
// with a zero-integral Lagrange row (uniform Riemann weights), which is equivalent to projecting the right-hand side onto
// zero mean, solving A p = b - mean(b), and returning the zero-mean solution.  The reference runs scipy BiCGStab with an
// ILU preconditioner on the CPU; here it is batched conjugate gradients on the GPU: three HBM-bound kernels per iteration,
// per-sample step lengths computed on the device (no host synchronisation inside the loop), dot products by warp shuffles
// + one double atomic per CTA, convergence per sample (a converged sample freezes: alpha = beta = 0).
// A is singular: rounding leaves a constant component in r that A can never reduce (p.Ap does not see it, r.r does), so
// alpha = r.r / p.Ap explodes once |r| approaches it (fp32: ~sqrt(n) eps |b| = 6e-6 |b| at 100 x 100, above a 1e-6 tolerance).
// Every use of r is therefore projected onto zero mean: sum(r) is accumulated next to r.r, p = (r - mean r) + beta p, and
// the norms are |r|^2 - (sum r)^2 / n.
#include <algorithm>

#include "pcnn_common.cuh"

namespace pcnn {

constexpr int KR_THREADS = 256;

__device__ __forceinline__ void block_atomic_add(double v, double* dst) {
    __shared__ double red[KR_THREADS / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < KR_THREADS / 32 ? red[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(dst, t);
    }
    __syncthreads();
}

// (A v)[i,j] = (deg * v[i,j] - sum of existing neighbours) / dh^2
__device__ __forceinline__ float neumann_apply_at(const float* __restrict__ v, int i, int j, int H, int W, float q) {
    const float c = v[(long long)i * W + j];
    float acc = 0.f;
    int deg = 0;
    if (i > 0) { acc += v[(long long)(i - 1) * W + j]; ++deg; }
    if (i < H - 1) { acc += v[(long long)(i + 1) * W + j]; ++deg; }
    if (j > 0) { acc += v[(long long)i * W + j - 1]; ++deg; }
    if (j < W - 1) { acc += v[(long long)i * W + j + 1]; ++deg; }
    return ((float)deg * c - acc) * q;
}

// sums[b] += sum x[b,:]   (grid (blocks, B))
__global__ void __launch_bounds__(KR_THREADS) kr_sum_kernel(const float* __restrict__ x, double* __restrict__ sums, long long n) {
    const int b = blockIdx.y;
    const float* xb = x + (long long)b * n;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)KR_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * KR_THREADS) acc += (double)xb[i];
    block_atomic_add(acc, sums + b);
}

// r = (-rhs - mean(-rhs)) - gs A x0,  p = 0 (the first direction kernel makes it r - mean r),  rr += r.r,  rs += sum r,
// bb += |b|^2   (x is scaled by gs afterwards)
__global__ void __launch_bounds__(KR_THREADS) kr_init_kernel(const float* __restrict__ rhs, const float* __restrict__ dx,
                                                            const float* __restrict__ guess_scale, const float* __restrict__ x,
                                                            float* __restrict__ r, float* __restrict__ p,
                                                            const double* __restrict__ rhs_sum, double* __restrict__ rr,
                                                            double* __restrict__ rs, double* __restrict__ bb, int H, int W) {
    const int b = blockIdx.y;
    const long long n = (long long)H * W;
    const float q = 1.0f / (dx[b] * dx[b]);
    const float mean_b = (float)(-rhs_sum[b] / (double)n);
    const float gs = guess_scale ? guess_scale[b] : 1.0f;
    const float* xb = x + b * n;
    const float* fb = rhs + b * n;
    double a_rr = 0.0, a_rs = 0.0, a_bb = 0.0;
    for (long long idx = blockIdx.x * (long long)KR_THREADS + threadIdx.x; idx < n; idx += (long long)gridDim.x * KR_THREADS) {
        const int i = (int)(idx / W), j = (int)(idx - (long long)i * W);
        const float bv = -fb[idx] - mean_b;
        const float rv = bv - gs * neumann_apply_at(xb, i, j, H, W, q);
        r[b * n + idx] = rv;
        p[b * n + idx] = 0.f;
        a_rr += (double)rv * rv;
        a_rs += (double)rv;
        a_bb += (double)bv * bv;
    }
    block_atomic_add(a_rr, rr + b);
    block_atomic_add(a_rs, rs + b);
    block_atomic_add(a_bb, bb + b);
}

__global__ void kr_scale_kernel(float* __restrict__ x, const float* __restrict__ scale, long long n, long long total) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        x[idx] *= scale[idx / n];
}

// q = A p ; pq[b] += p.q ; zeroes the accumulators the NEXT kernel will add into
__global__ void __launch_bounds__(KR_THREADS) kr_apply_kernel(const float* __restrict__ p, const float* __restrict__ dx,
                                                             float* __restrict__ qv, double* __restrict__ pq,
                                                             double* __restrict__ rr_next, double* __restrict__ rs_next, int H, int W) {
    const int b = blockIdx.y;
    const long long n = (long long)H * W;
    if (blockIdx.x == 0 && threadIdx.x == 0) { rr_next[b] = 0.0; rs_next[b] = 0.0; }
    const float q = 1.0f / (dx[b] * dx[b]);
    const float* pb = p + b * n;
    double acc = 0.0;
    for (long long idx = blockIdx.x * (long long)KR_THREADS + threadIdx.x; idx < n; idx += (long long)gridDim.x * KR_THREADS) {
        const int i = (int)(idx / W), j = (int)(idx - (long long)i * W);
        const float v = neumann_apply_at(pb, i, j, H, W, q);
        qv[b * n + idx] = v;
        acc += (double)pb[idx] * v;
    }
    block_atomic_add(acc, pq + b);
}

// alpha = |r - mean r|^2 / p.q (0 once converged) ; x += alpha p ; r -= alpha q ; rr_next += r.r ; rs_next += sum r
__global__ void __launch_bounds__(KR_THREADS) kr_update_kernel(float* __restrict__ x, float* __restrict__ r,
                                                              const float* __restrict__ p, const float* __restrict__ qv,
                                                              const double* __restrict__ rr, const double* __restrict__ rs,
                                                              const double* __restrict__ pq, const double* __restrict__ bb,
                                                              double* __restrict__ rr_next, double* __restrict__ rs_next,
                                                              double tol2, long long n) {
    const int b = blockIdx.y;
    const double rrp = fmax(rr[b] - rs[b] * rs[b] / (double)n, 0.0);
    const bool active = rrp > tol2 * bb[b] && pq[b] > 0.0;
    const float alpha = active ? (float)(rrp / pq[b]) : 0.f;
    double acc = 0.0, accs = 0.0;
    for (long long idx = blockIdx.x * (long long)KR_THREADS + threadIdx.x; idx < n; idx += (long long)gridDim.x * KR_THREADS) {
        const long long g = b * n + idx;
        x[g] = fmaf(alpha, p[g], x[g]);
        const float rv = fmaf(-alpha, qv[g], r[g]);
        r[g] = rv;
        acc += (double)rv * rv;
        accs += (double)rv;
    }
    block_atomic_add(acc, rr_next + b);
    block_atomic_add(accs, rs_next + b);
}

// beta = |r_next|^2 / |r|^2 (projected norms; 0 once converged and for the very first direction) ; p = (r - mean r) + beta p ;
// records the relative residual ; zeroes pq for the next iteration
__global__ void __launch_bounds__(KR_THREADS) kr_direction_kernel(float* __restrict__ p, const float* __restrict__ r,
                                                                 const double* __restrict__ rr, const double* __restrict__ rs,
                                                                 const double* __restrict__ rr_next, const double* __restrict__ rs_next,
                                                                 const double* __restrict__ bb, double* __restrict__ pq,
                                                                 double* __restrict__ history, double tol2, long long n, int first) {
    const int b = blockIdx.y;
    const double rrp = fmax(rr[b] - rs[b] * rs[b] / (double)n, 0.0);
    const double rrp_next = fmax(rr_next[b] - rs_next[b] * rs_next[b] / (double)n, 0.0);
    const bool active = !first && rrp > tol2 * bb[b];
    const float beta = active ? (float)(rrp_next / rrp) : 0.f;
    const float mean = (float)(rs_next[b] / (double)n);
    for (long long idx = blockIdx.x * (long long)KR_THREADS + threadIdx.x; idx < n; idx += (long long)gridDim.x * KR_THREADS) {
        const long long g = b * n + idx;
        p[g] = fmaf(beta, p[g], r[g] - mean);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (history) history[b] = bb[b] > 0.0 ? sqrt(rrp_next / bb[b]) : 0.0;
        pq[b] = 0.0;      // no block of this kernel reads pq (kr_update did); the next kr_apply accumulates into it
    }
}

// ---- float4 variants of the three per-iteration kernels (W % 4 == 0, 16-byte aligned vectors): one 16-byte access per
// vector and thread instead of four 4-byte ones, one 32-bit division per 4 points instead of a 64-bit one per point.
// The arithmetic per point follows the scalar kernels (same order of the neighbour sum).

// 4 points (i, 4*j4 .. 4*j4+3) of A v.  v4: the sample's map as float4 rows of W4 = W/4
__device__ __forceinline__ float4 neumann_apply_vec4(const float4* __restrict__ v4, const float* __restrict__ v, int i, int j4, int H, int W4,
                                                     float q, float4& centre) {
    const long long row = (long long)i * W4;
    const float4 c = v4[row + j4];
    centre = c;
    const bool has_up = i > 0, has_dn = i < H - 1, has_l = j4 > 0, has_r = j4 < W4 - 1;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 up = has_up ? v4[row - W4 + j4] : z;
    const float4 dn = has_dn ? v4[row + W4 + j4] : z;
    const float left = has_l ? v[(row + j4) * 4 - 1] : 0.f;
    const float right = has_r ? v[(row + j4) * 4 + 4] : 0.f;
    const int dv = (int)has_up + (int)has_dn;
    float4 o;
    // scalar order: ((up + down) + left) + right, skipping the neighbours that do not exist (adding 0.f changes nothing)
    o.x = ((float)(dv + 1 + (int)has_l) * c.x - (((up.x + dn.x) + left) + c.y)) * q;
    o.y = ((float)(dv + 2) * c.y - (((up.y + dn.y) + c.x) + c.z)) * q;
    o.z = ((float)(dv + 2) * c.z - (((up.z + dn.z) + c.y) + c.w)) * q;
    o.w = ((float)(dv + 1 + (int)has_r) * c.w - (((up.w + dn.w) + c.z) + right)) * q;
    return o;
}

__global__ void __launch_bounds__(KR_THREADS) kr_apply_vec4_kernel(const float* __restrict__ p, const float* __restrict__ dx,
                                                                  float* __restrict__ qv, double* __restrict__ pq,
                                                                  double* __restrict__ rr_next, double* __restrict__ rs_next, int H, int W4) {
    const int b = blockIdx.y;
    const unsigned n4 = (unsigned)H * (unsigned)W4;
    if (blockIdx.x == 0 && threadIdx.x == 0) { rr_next[b] = 0.0; rs_next[b] = 0.0; }
    const float q = 1.0f / (dx[b] * dx[b]);
    const float* pb = p + (long long)b * n4 * 4;
    const float4* pb4 = reinterpret_cast<const float4*>(pb);
    float4* qb4 = reinterpret_cast<float4*>(qv + (long long)b * n4 * 4);
    double acc = 0.0;
    for (unsigned v = blockIdx.x * KR_THREADS + threadIdx.x; v < n4; v += gridDim.x * KR_THREADS) {
        const unsigned i = v / (unsigned)W4, j4 = v - i * (unsigned)W4;
        float4 c;
        const float4 o = neumann_apply_vec4(pb4, pb, (int)i, (int)j4, H, W4, q, c);
        qb4[v] = o;
        acc += (double)c.x * o.x;
        acc += (double)c.y * o.y;
        acc += (double)c.z * o.z;
        acc += (double)c.w * o.w;
    }
    block_atomic_add(acc, pq + b);
}

__global__ void __launch_bounds__(KR_THREADS) kr_update_vec4_kernel(float* __restrict__ x, float* __restrict__ r,
                                                                   const float* __restrict__ p, const float* __restrict__ qv,
                                                                   const double* __restrict__ rr, const double* __restrict__ rs,
                                                                   const double* __restrict__ pq, const double* __restrict__ bb,
                                                                   double* __restrict__ rr_next, double* __restrict__ rs_next,
                                                                   double tol2, long long n) {
    const int b = blockIdx.y;
    const double rrp = fmax(rr[b] - rs[b] * rs[b] / (double)n, 0.0);
    const bool active = rrp > tol2 * bb[b] && pq[b] > 0.0;
    const float alpha = active ? (float)(rrp / pq[b]) : 0.f;
    const unsigned n4 = (unsigned)(n >> 2);
    float4* x4 = reinterpret_cast<float4*>(x + b * n);
    float4* r4 = reinterpret_cast<float4*>(r + b * n);
    const float4* p4 = reinterpret_cast<const float4*>(p + b * n);
    const float4* q4 = reinterpret_cast<const float4*>(qv + b * n);
    double acc = 0.0, accs = 0.0;
    for (unsigned v = blockIdx.x * KR_THREADS + threadIdx.x; v < n4; v += gridDim.x * KR_THREADS) {
        float4 xv = x4[v], rv = r4[v];
        const float4 pv = p4[v], qq = q4[v];
        xv.x = fmaf(alpha, pv.x, xv.x); xv.y = fmaf(alpha, pv.y, xv.y); xv.z = fmaf(alpha, pv.z, xv.z); xv.w = fmaf(alpha, pv.w, xv.w);
        rv.x = fmaf(-alpha, qq.x, rv.x); rv.y = fmaf(-alpha, qq.y, rv.y); rv.z = fmaf(-alpha, qq.z, rv.z); rv.w = fmaf(-alpha, qq.w, rv.w);
        x4[v] = xv;
        r4[v] = rv;
        acc += (double)rv.x * rv.x; acc += (double)rv.y * rv.y; acc += (double)rv.z * rv.z; acc += (double)rv.w * rv.w;
        accs += (double)rv.x; accs += (double)rv.y; accs += (double)rv.z; accs += (double)rv.w;
    }
    block_atomic_add(acc, rr_next + b);
    block_atomic_add(accs, rs_next + b);
}

__global__ void __launch_bounds__(KR_THREADS) kr_direction_vec4_kernel(float* __restrict__ p, const float* __restrict__ r,
                                                                      const double* __restrict__ rr, const double* __restrict__ rs,
                                                                      const double* __restrict__ rr_next, const double* __restrict__ rs_next,
                                                                      const double* __restrict__ bb, double* __restrict__ pq,
                                                                      double* __restrict__ history, double tol2, long long n, int first) {
    const int b = blockIdx.y;
    const double rrp = fmax(rr[b] - rs[b] * rs[b] / (double)n, 0.0);
    const double rrp_next = fmax(rr_next[b] - rs_next[b] * rs_next[b] / (double)n, 0.0);
    const bool active = !first && rrp > tol2 * bb[b];
    const float beta = active ? (float)(rrp_next / rrp) : 0.f;
    const float mean = (float)(rs_next[b] / (double)n);
    const unsigned n4 = (unsigned)(n >> 2);
    float4* p4 = reinterpret_cast<float4*>(p + b * n);
    const float4* r4 = reinterpret_cast<const float4*>(r + b * n);
    for (unsigned v = blockIdx.x * KR_THREADS + threadIdx.x; v < n4; v += gridDim.x * KR_THREADS) {
        float4 pv = p4[v];
        const float4 rv = r4[v];
        pv.x = fmaf(beta, pv.x, rv.x - mean); pv.y = fmaf(beta, pv.y, rv.y - mean);
        pv.z = fmaf(beta, pv.z, rv.z - mean); pv.w = fmaf(beta, pv.w, rv.w - mean);
        p4[v] = pv;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (history) history[b] = bb[b] > 0.0 ? sqrt(rrp_next / bb[b]) : 0.0;
        pq[b] = 0.0;
    }
}

// out = A v (the operator alone: tests, residual checks)
__global__ void __launch_bounds__(KR_THREADS) kr_apply_only_kernel(const float* __restrict__ p, const float* __restrict__ dx,
                                                                  float* __restrict__ q, int H, int W) {
    const int b = blockIdx.y;
    const long long n = (long long)H * W;
    const float qq = 1.0f / (dx[b] * dx[b]);
    for (long long idx = blockIdx.x * (long long)KR_THREADS + threadIdx.x; idx < n; idx += (long long)gridDim.x * KR_THREADS) {
        const int i = (int)(idx / W), j = (int)(idx - (long long)i * W);
        q[b * n + idx] = neumann_apply_at(p + b * n, i, j, H, W, qq);
    }
}

// x -= mean(x)
__global__ void kr_center_kernel(float* __restrict__ x, const double* __restrict__ sums, long long n, long long total) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        x[idx] -= (float)(sums[idx / n] / (double)n);
}

}  // namespace pcnn

using namespace pcnn;

extern "C" size_t pcnn_neumann_cg_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H < 2 || W < 2) return 0;
    // r, p, q (float, B*H*W each) + 6 double accumulators per sample
    return (size_t)3 * B * H * W * sizeof(float) + (size_t)8 * B * sizeof(double);
}

extern "C" int pcnn_neumann_laplacian_apply_f32(const float* v, const float* dx, float* out, int B, int H, int W, void* stream) {
    PCNN_CHECK_ARG(v && dx && out && v != out && B > 0 && B <= 65535 && H >= 2 && W >= 2, "neumann_laplacian_apply_f32: bad argument");
    const long long n = (long long)H * W;
    const int gx = (int)std::min<long long>((n + KR_THREADS * 4 - 1) / (KR_THREADS * 4), 2048);
    kr_apply_only_kernel<<<dim3(gx, B), KR_THREADS, 0, (cudaStream_t)stream>>>(v, dx, out, H, W);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

extern "C" int pcnn_neumann_cg_solve(const float* rhs, const float* dx, const float* guess_scale, float* x, int B, int H, int W,
                                     int max_iter, double rel_tol, double* residual_history, void* workspace, void* stream) {
    PCNN_CHECK_ARG(rhs && dx && x && workspace, "neumann_cg_solve: null pointer");
    PCNN_CHECK_ARG(B > 0 && B <= 65535 && H >= 2 && W >= 2 && max_iter >= 0 && rel_tol >= 0.0, "neumann_cg_solve: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)H * W, total = n * B;
    float* r = reinterpret_cast<float*>(workspace);
    float* p = r + total;
    float* q = p + total;
    double* acc = reinterpret_cast<double*>(q + total);      // [rhs_sum | bb | pq | rr0 | rr1 | xsum | rs0 | rs1] x B
    PCNN_CHECK_ARG(((uintptr_t)acc & 7) == 0, "neumann_cg_solve: workspace must be 8-byte aligned");
    double *rhs_sum = acc, *bb = acc + B, *pq = acc + 2 * B, *rr0 = acc + 3 * B, *rr1 = acc + 4 * B, *xsum = acc + 5 * B;
    double *rs0 = acc + 6 * B, *rs1 = acc + 7 * B;
    PCNN_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 8 * B, st));
    const int gx = (int)std::min<long long>((n + KR_THREADS * 4 - 1) / (KR_THREADS * 4), 1024);
    const dim3 grid(gx, B);
    const double tol2 = rel_tol * rel_tol;
    // float4 kernels when every row is a whole number of 16-byte vectors (r, p, q start n*B floats apart: aligned with x and the workspace)
    const bool vec4 = (W % 4 == 0) && (((uintptr_t)x | (uintptr_t)workspace) & 15) == 0 && n / 4 < (1ll << 31);
    kr_sum_kernel<<<grid, KR_THREADS, 0, st>>>(rhs, rhs_sum, n);
    PCNN_CHECK_LAUNCH();
    kr_init_kernel<<<grid, KR_THREADS, 0, st>>>(rhs, dx, guess_scale, x, r, p, rhs_sum, rr0, rs0, bb, H, W);
    PCNN_CHECK_LAUNCH();
    if (guess_scale) {
        const int gs = (int)std::min<long long>((total + 255) / 256, 148 * 16);
        kr_scale_kernel<<<gs, 256, 0, st>>>(x, guess_scale, n, total);
        PCNN_CHECK_LAUNCH();
    }
    kr_direction_kernel<<<grid, KR_THREADS, 0, st>>>(p, r, rr0, rs0, rr0, rs0, bb, pq, nullptr, tol2, n, 1);     // p = r - mean(r)
    PCNN_CHECK_LAUNCH();
    for (int it = 0; it < max_iter; ++it) {
        double *rr = (it & 1) ? rr1 : rr0, *rs = (it & 1) ? rs1 : rs0;
        double *rr_next = (it & 1) ? rr0 : rr1, *rs_next = (it & 1) ? rs0 : rs1;
        double* hist = residual_history ? residual_history + (size_t)it * B : nullptr;
        if (vec4) {
            kr_apply_vec4_kernel<<<grid, KR_THREADS, 0, st>>>(p, dx, q, pq, rr_next, rs_next, H, W / 4);
            PCNN_CHECK_LAUNCH();
            kr_update_vec4_kernel<<<grid, KR_THREADS, 0, st>>>(x, r, p, q, rr, rs, pq, bb, rr_next, rs_next, tol2, n);
            PCNN_CHECK_LAUNCH();
            kr_direction_vec4_kernel<<<grid, KR_THREADS, 0, st>>>(p, r, rr, rs, rr_next, rs_next, bb, pq, hist, tol2, n, 0);
            PCNN_CHECK_LAUNCH();
            continue;
        }
        kr_apply_kernel<<<grid, KR_THREADS, 0, st>>>(p, dx, q, pq, rr_next, rs_next, H, W);
        PCNN_CHECK_LAUNCH();
        kr_update_kernel<<<grid, KR_THREADS, 0, st>>>(x, r, p, q, rr, rs, pq, bb, rr_next, rs_next, tol2, n);
        PCNN_CHECK_LAUNCH();
        kr_direction_kernel<<<grid, KR_THREADS, 0, st>>>(p, r, rr, rs, rr_next, rs_next, bb, pq, hist, tol2, n, 0);
        PCNN_CHECK_LAUNCH();
    }
    kr_sum_kernel<<<grid, KR_THREADS, 0, st>>>(x, xsum, n);
    PCNN_CHECK_LAUNCH();
    const int gs = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    kr_center_kernel<<<gs, 256, 0, st>>>(x, xsum, n, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
