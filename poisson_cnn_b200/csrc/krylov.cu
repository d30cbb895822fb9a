// The step AFTER the hot path: the pressure-Poisson solve of the reference's Navier-Stokes projection
// solver, which takes the Neumann HPNN's prediction as the initial guess of a Krylov iteration
// (Navier_Stokes_2D/solvers.py:29-33 builds the model, :153-186 the operator, :204-334 the solve).
//
// Operator (Poisson_pressure_matrix, solvers.py:164-173): A = kron(I, T) + kron(T, I), T = tridiag(-1, [1,2,...,2,1], -1)/dh^2,
// i.e. minus the cell-centred 5-point Laplacian with homogeneous Neumann boundaries (the diagonal counts the neighbours
// that exist).  A is symmetric positive semi-definite with the constants as null space; the reference pins the constant
// with a zero-integral Lagrange row (uniform Riemann weights), which is equivalent to projecting the right-hand side onto
// zero mean, solving A p = b - mean(b), and returning the zero-mean solution.  The reference runs scipy BiCGStab with an
// ILU preconditioner on the CPU; here it is batched conjugate gradients on the GPU: two HBM-bound kernels per iteration,
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

// ---- one CG iteration = TWO kernels, 36 B of fp32 vectors per grid point:
//   kr_direction_apply:  p_k = (r_k - mean r_k) + beta p_{k-1}   (written to the OTHER p buffer: neighbours still need p_{k-1}),
//                        x  += alpha_{k-1} p_{k-1}               (the previous iteration's update of x, deferred to here
//                                                                 because p_{k-1} is being read anyway),
//                        q   = A p_k  (p_k of the four neighbours is recomputed from their r and p_{k-1}),  pq_k += p_k.q
//                        reads r p x (+ halo from L1/L2), writes p x q: 24 B
//   kr_residual:         r  -= alpha_k q ;  rr_{k+1} += r.r ;  rs_{k+1} += sum r        reads r q, writes r: 12 B
// alpha_k = |r_k - mean|^2 / pq_k and beta = |r_k - mean|^2 / |r_{k-1} - mean|^2 are recomputed by every block from the
// per-sample double accumulators, which rotate (3 generations of rr / rs, 2 of pq) so that a kernel never zeroes or
// accumulates into a slot another block of the same kernel reads.  Frozen (converged) samples get alpha = beta = 0.
// V4: one thread per 4 consecutive points of a row (W % 4 == 0, 16-byte aligned vectors), else one thread per point.

__device__ __forceinline__ double projected_norm2(const double* rr, const double* rs, int b, long long n) {
    return fmax(rr[b] - rs[b] * rs[b] / (double)n, 0.0);
}

template <bool V4>
__global__ void __launch_bounds__(KR_THREADS) kr_direction_apply_kernel(
        const float* __restrict__ r, const float* __restrict__ p_old, float* __restrict__ p_new, float* __restrict__ x,
        float* __restrict__ qv, const float* __restrict__ dx, const double* __restrict__ bb,
        const double* __restrict__ rr_prev, const double* __restrict__ rs_prev, const double* __restrict__ pq_prev,
        const double* __restrict__ rr_cur, const double* __restrict__ rs_cur, double* __restrict__ pq_cur,
        double* __restrict__ rr_zero, double* __restrict__ rs_zero, double* __restrict__ history, double tol2,
        int H, int Wv, int first) {
    const int b = blockIdx.y;
    const unsigned nv = (unsigned)H * (unsigned)Wv;
    const long long n = (long long)nv * (V4 ? 4 : 1);
    const double rrp_cur = projected_norm2(rr_cur, rs_cur, b, n);
    float alpha_prev = 0.f, beta = 0.f;
    if (!first) {
        const double rrp_prev = projected_norm2(rr_prev, rs_prev, b, n);
        const bool active = rrp_prev > tol2 * bb[b];
        if (active) beta = (float)(rrp_cur / rrp_prev);
        if (active && pq_prev[b] > 0.0) alpha_prev = (float)(rrp_prev / pq_prev[b]);
    }
    const float mean = (float)(rs_cur[b] / (double)n);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rr_zero[b] = 0.0;
        rs_zero[b] = 0.0;
        if (history) history[b] = bb[b] > 0.0 ? sqrt(rrp_cur / bb[b]) : 0.0;
    }
    const float q = 1.0f / (dx[b] * dx[b]);
    const float* rb = r + b * n;
    const float* pb = p_old + b * n;
    double acc = 0.0;
    for (unsigned v = blockIdx.x * KR_THREADS + threadIdx.x; v < nv; v += gridDim.x * KR_THREADS) {
        const unsigned i = v / (unsigned)Wv, j = v - i * (unsigned)Wv;
        const bool has_up = i > 0, has_dn = i < (unsigned)H - 1, has_l = j > 0, has_r = j < (unsigned)Wv - 1;
        const int dv = (int)has_up + (int)has_dn;
        if constexpr (V4) {
            const float4* r4 = reinterpret_cast<const float4*>(rb);
            const float4* p4 = reinterpret_cast<const float4*>(pb);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 cr = r4[v], cp = p4[v];
            const float4 ur = has_up ? r4[v - Wv] : z, upo = has_up ? p4[v - Wv] : z;
            const float4 dr = has_dn ? r4[v + Wv] : z, dpo = has_dn ? p4[v + Wv] : z;
            float4 c, up = z, dn = z;
            c.x = fmaf(beta, cp.x, cr.x - mean); c.y = fmaf(beta, cp.y, cr.y - mean);
            c.z = fmaf(beta, cp.z, cr.z - mean); c.w = fmaf(beta, cp.w, cr.w - mean);
            if (has_up) {
                up.x = fmaf(beta, upo.x, ur.x - mean); up.y = fmaf(beta, upo.y, ur.y - mean);
                up.z = fmaf(beta, upo.z, ur.z - mean); up.w = fmaf(beta, upo.w, ur.w - mean);
            }
            if (has_dn) {
                dn.x = fmaf(beta, dpo.x, dr.x - mean); dn.y = fmaf(beta, dpo.y, dr.y - mean);
                dn.z = fmaf(beta, dpo.z, dr.z - mean); dn.w = fmaf(beta, dpo.w, dr.w - mean);
            }
            const long long e = (long long)v * 4;
            const float left = has_l ? fmaf(beta, pb[e - 1], rb[e - 1] - mean) : 0.f;
            const float right = has_r ? fmaf(beta, pb[e + 4], rb[e + 4] - mean) : 0.f;
            float4 o;      // the scalar operator's order: ((up + down) + left) + right
            o.x = ((float)(dv + 1 + (int)has_l) * c.x - (((up.x + dn.x) + left) + c.y)) * q;
            o.y = ((float)(dv + 2) * c.y - (((up.y + dn.y) + c.x) + c.z)) * q;
            o.z = ((float)(dv + 2) * c.z - (((up.z + dn.z) + c.y) + c.w)) * q;
            o.w = ((float)(dv + 1 + (int)has_r) * c.w - (((up.w + dn.w) + c.z) + right)) * q;
            float4* x4 = reinterpret_cast<float4*>(x + b * n);
            float4 xv = x4[v];
            xv.x = fmaf(alpha_prev, cp.x, xv.x); xv.y = fmaf(alpha_prev, cp.y, xv.y);
            xv.z = fmaf(alpha_prev, cp.z, xv.z); xv.w = fmaf(alpha_prev, cp.w, xv.w);
            x4[v] = xv;
            reinterpret_cast<float4*>(p_new + b * n)[v] = c;
            reinterpret_cast<float4*>(qv + b * n)[v] = o;
            acc += (double)c.x * o.x; acc += (double)c.y * o.y; acc += (double)c.z * o.z; acc += (double)c.w * o.w;
        } else {
            const float cp = pb[v];
            const float c = fmaf(beta, cp, rb[v] - mean);
            float nb = 0.f;
            if (has_up) nb += fmaf(beta, pb[v - Wv], rb[v - Wv] - mean);
            if (has_dn) nb += fmaf(beta, pb[v + Wv], rb[v + Wv] - mean);
            if (has_l) nb += fmaf(beta, pb[v - 1], rb[v - 1] - mean);
            if (has_r) nb += fmaf(beta, pb[v + 1], rb[v + 1] - mean);
            const float o = ((float)(dv + (int)has_l + (int)has_r) * c - nb) * q;
            x[b * n + v] = fmaf(alpha_prev, cp, x[b * n + v]);
            p_new[b * n + v] = c;
            qv[b * n + v] = o;
            acc += (double)c * o;
        }
    }
    block_atomic_add(acc, pq_cur + b);
}

template <bool V4>
__global__ void __launch_bounds__(KR_THREADS) kr_residual_kernel(float* __restrict__ r, const float* __restrict__ qv,
                                                                const double* __restrict__ rr_cur, const double* __restrict__ rs_cur,
                                                                const double* __restrict__ pq_cur, const double* __restrict__ bb,
                                                                double* __restrict__ rr_next, double* __restrict__ rs_next,
                                                                double* __restrict__ pq_zero, double tol2, long long n) {
    const int b = blockIdx.y;
    const double rrp = projected_norm2(rr_cur, rs_cur, b, n);
    const bool active = rrp > tol2 * bb[b] && pq_cur[b] > 0.0;
    const float alpha = active ? (float)(rrp / pq_cur[b]) : 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) pq_zero[b] = 0.0;
    double acc = 0.0, accs = 0.0;
    if constexpr (V4) {
        const unsigned n4 = (unsigned)(n >> 2);
        float4* r4 = reinterpret_cast<float4*>(r + b * n);
        const float4* q4 = reinterpret_cast<const float4*>(qv + b * n);
        for (unsigned v = blockIdx.x * KR_THREADS + threadIdx.x; v < n4; v += gridDim.x * KR_THREADS) {
            float4 rv = r4[v];
            const float4 qq = q4[v];
            rv.x = fmaf(-alpha, qq.x, rv.x); rv.y = fmaf(-alpha, qq.y, rv.y); rv.z = fmaf(-alpha, qq.z, rv.z); rv.w = fmaf(-alpha, qq.w, rv.w);
            r4[v] = rv;
            acc += (double)rv.x * rv.x; acc += (double)rv.y * rv.y; acc += (double)rv.z * rv.z; acc += (double)rv.w * rv.w;
            accs += (double)rv.x; accs += (double)rv.y; accs += (double)rv.z; accs += (double)rv.w;
        }
    } else {
        for (long long idx = blockIdx.x * (long long)KR_THREADS + threadIdx.x; idx < n; idx += (long long)gridDim.x * KR_THREADS) {
            const float rv = fmaf(-alpha, qv[b * n + idx], r[b * n + idx]);
            r[b * n + idx] = rv;
            acc += (double)rv * rv;
            accs += (double)rv;
        }
    }
    block_atomic_add(acc, rr_next + b);
    block_atomic_add(accs, rs_next + b);
}

// after the last iteration: the deferred x += alpha_last p_last, and the last row of the residual history
__global__ void __launch_bounds__(KR_THREADS) kr_final_kernel(float* __restrict__ x, const float* __restrict__ p_last,
                                                             const double* __restrict__ rr_last, const double* __restrict__ rs_last,
                                                             const double* __restrict__ pq_last, const double* __restrict__ rr_end,
                                                             const double* __restrict__ rs_end, const double* __restrict__ bb,
                                                             double* __restrict__ history, double tol2, long long n) {
    const int b = blockIdx.y;
    const double rrp = projected_norm2(rr_last, rs_last, b, n);
    const bool active = rrp > tol2 * bb[b] && pq_last[b] > 0.0;
    const float alpha = active ? (float)(rrp / pq_last[b]) : 0.f;
    if (history && blockIdx.x == 0 && threadIdx.x == 0)
        history[b] = bb[b] > 0.0 ? sqrt(projected_norm2(rr_end, rs_end, b, n) / bb[b]) : 0.0;
    for (long long idx = blockIdx.x * (long long)KR_THREADS + threadIdx.x; idx < n; idx += (long long)gridDim.x * KR_THREADS)
        x[b * n + idx] = fmaf(alpha, p_last[b * n + idx], x[b * n + idx]);
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
    // r, two p buffers, q (float, B*H*W each) + 12 double accumulators per sample
    return (size_t)4 * B * H * W * sizeof(float) + (size_t)12 * B * sizeof(double);
}

extern "C" int pcnn_neumann_laplacian_apply_f32(const float* v, const float* dx, float* out, int B, int H, int W, void* stream) {
    PCNN_CHECK_ARG(v && dx && out && v != out && B > 0 && B <= 65535 && H >= 2 && W >= 2, "neumann_laplacian_apply_f32: bad argument");
    const long long n = (long long)H * W;
    const int gx = (int)std::min<long long>((n + KR_THREADS * 4 - 1) / (KR_THREADS * 4), 2048);
    kr_apply_only_kernel<<<dim3(gx, B), KR_THREADS, 0, (cudaStream_t)stream>>>(v, dx, out, H, W);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

template <bool V4>
static int cg_iterations(float* x, float* r, float* pbuf0, float* pbuf1, float* q, const float* dx, const double* bb, double* const* rr,
                         double* const* rs, double* const* pq, int B, int H, int W, int max_iter, double tol2, double* residual_history,
                         dim3 grid, cudaStream_t st) {
    const long long n = (long long)H * W;
    float* pbuf[2] = {pbuf0, pbuf1};
    for (int k = 0; k < max_iter; ++k) {
        const int c3 = k % 3, p3 = (k + 2) % 3, n3 = (k + 1) % 3, c2 = k & 1, o2 = (k + 1) & 1;
        kr_direction_apply_kernel<V4><<<grid, KR_THREADS, 0, st>>>(r, pbuf[c2], pbuf[o2], x, q, dx, bb, rr[p3], rs[p3], pq[o2], rr[c3], rs[c3], pq[c2],
                                                                   rr[n3], rs[n3], (residual_history && k > 0) ? residual_history + (size_t)(k - 1) * B : nullptr,
                                                                   tol2, H, V4 ? W / 4 : W, k == 0);
        PCNN_CHECK_LAUNCH();
        kr_residual_kernel<V4><<<grid, KR_THREADS, 0, st>>>(r, q, rr[c3], rs[c3], pq[c2], bb, rr[n3], rs[n3], pq[o2], tol2, n);
        PCNN_CHECK_LAUNCH();
    }
    if (max_iter > 0) {
        const int k = max_iter - 1;
        kr_final_kernel<<<grid, KR_THREADS, 0, st>>>(x, pbuf[(k + 1) & 1], rr[k % 3], rs[k % 3], pq[k & 1], rr[(k + 1) % 3], rs[(k + 1) % 3], bb,
                                                     residual_history ? residual_history + (size_t)k * B : nullptr, tol2, n);
        PCNN_CHECK_LAUNCH();
    }
    return PCNN_OK;
}

extern "C" int pcnn_neumann_cg_solve(const float* rhs, const float* dx, const float* guess_scale, float* x, int B, int H, int W,
                                     int max_iter, double rel_tol, double* residual_history, void* workspace, void* stream) {
    PCNN_CHECK_ARG(rhs && dx && x && workspace, "neumann_cg_solve: null pointer");
    PCNN_CHECK_ARG(B > 0 && B <= 65535 && H >= 2 && W >= 2 && max_iter >= 0 && rel_tol >= 0.0, "neumann_cg_solve: bad argument");
    PCNN_CHECK_ARG((long long)H * W < (1ll << 31), "neumann_cg_solve: at most 2^31 - 1 points per sample");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)H * W, total = n * B;
    float* r = reinterpret_cast<float*>(workspace);
    float* p0 = r + total;
    float* p1 = p0 + total;
    float* q = p1 + total;
    double* acc = reinterpret_cast<double*>(q + total);      // [rhs_sum | bb | xsum | pq x 2 | rr x 3 | rs x 3] x B
    PCNN_CHECK_ARG(((uintptr_t)acc & 7) == 0, "neumann_cg_solve: workspace must be 8-byte aligned");
    double *rhs_sum = acc, *bb = acc + B, *xsum = acc + 2 * B;
    double* pq[2] = {acc + 3 * B, acc + 4 * B};
    double* rr[3] = {acc + 5 * B, acc + 6 * B, acc + 7 * B};
    double* rs[3] = {acc + 8 * B, acc + 9 * B, acc + 10 * B};
    PCNN_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 12 * B, st));
    const int gx = (int)std::min<long long>((n + KR_THREADS * 4 - 1) / (KR_THREADS * 4), 1024);
    const dim3 grid(gx, B);
    const double tol2 = rel_tol * rel_tol;
    kr_sum_kernel<<<grid, KR_THREADS, 0, st>>>(rhs, rhs_sum, n);
    PCNN_CHECK_LAUNCH();
    kr_init_kernel<<<grid, KR_THREADS, 0, st>>>(rhs, dx, guess_scale, x, r, p0, rhs_sum, rr[0], rs[0], bb, H, W);     // p_{-1} = 0
    PCNN_CHECK_LAUNCH();
    if (guess_scale) {
        const int gs = (int)std::min<long long>((total + 255) / 256, 148 * 16);
        kr_scale_kernel<<<gs, 256, 0, st>>>(x, guess_scale, n, total);
        PCNN_CHECK_LAUNCH();
    }
    // float4 kernels when every row is a whole number of 16-byte vectors (r, p, q start n*B floats apart: aligned with x and the workspace)
    const bool vec4 = (W % 4 == 0) && (((uintptr_t)x | (uintptr_t)workspace) & 15) == 0;
    const int rc = vec4 ? cg_iterations<true>(x, r, p0, p1, q, dx, bb, rr, rs, pq, B, H, W, max_iter, tol2, residual_history, grid, st)
                        : cg_iterations<false>(x, r, p0, p1, q, dx, bb, rr, rs, pq, B, H, W, max_iter, tol2, residual_history, grid, st);
    if (rc != PCNN_OK) return rc;
    kr_sum_kernel<<<grid, KR_THREADS, 0, st>>>(x, xsum, n);
    PCNN_CHECK_LAUNCH();
    const int gs = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    kr_center_kernel<<<gs, 256, 0, st>>>(x, xsum, n, total);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}
