// HBM-bound DST-I direct Poisson solve (ground truth for accuracy checks).
//
// Solves the reference's discrete system (dataset/solvers/multigrid.py:98-150, dataset/solvers/cholesky.py:45-119)
//     (u[i+1,j] + u[i-1,j] + u[i,j+1] + u[i,j-1] - 4 u[i,j]) / dx^2 = f[i,j]   on the interior, Dirichlet ring,
// by diagonalising the 5-point operator with sine transforms:  u = DST_x DST_y [ DST_x DST_y b / (lam_x + lam_y) ] * norm.
//
// The interior has n = N - 2 points per side, so DST-I needs a length-2(n+1) = 2(N-1) transform: 510 = 2*3*5*17 at N = 256,
// 4094 = 2*23*89 at N = 2048 -- never a power of two for the grid sizes of BASELINE.json.  Every line is therefore transformed
// with Bluestein's chirp-z identity  jk = (j^2 + k^2 - (k-j)^2)/2 :
//     S_k = sum_j x_j sin(pi j k / M) = Im[ c_k * sum_j (x_j c_j) conj(c_{k-j}) ],   c_m = exp(i pi m^2 / (2M)),  M = n + 1,
// i.e. one circular convolution of length L = 2^q >= 2n - 1, done with two power-of-two FFTs in SHARED MEMORY (the spectrum of
// conj(c) is precomputed per length).  FFT: Stockham autosort, radix 8 (last stage radix 2 or 4), L/8 threads per transform
// with 8 points in registers per thread, in place (all reads, barrier, all writes), index padding idx + idx/8 makes both the
// strided reads and the scattered writes bank-conflict-free for 8- and 16-byte elements.
//
// Passes over the grid (the O(N^3) sine-matrix GEMM form in checks.cu made 4 passes of dense GEMMs):
//   P1 rows   : b = -dx^2 f + adjacent Dirichlet values (built on the fly), DST along y        f (4 B) -> T (8 B)
//   P2 columns: after the y-transform the system decouples per y-mode k into constant-coefficient tridiagonal systems
//               -u[i-1] + (2 + lam_k) u[i] - u[i+1] = T[i][k] along x: Thomas algorithm, one thread per (sample, k),
//               coalesced across k, float64 (condition ~1e6 for the lowest modes) -- "Fourier analysis + tridiagonal
//               solve" (FACR(0)); two sweeps over T in place instead of two more DSTs per column       T -> T (2 x 16 B)
//   P3 rows   : DST along y, write u (fp32) and the Dirichlet ring                               T (8 B) -> u (4 B)
// The FFT arithmetic is double (default: the reference solves in float64) or float; T and the tridiagonal solve are always
// double.  Algorithmic bytes: 8 B per grid point; actual traffic 8 + 16 + 32 = 56 B per point plus the L2-resident table
// of elimination coefficients, in 4 kernel launches ("passes": 3 full-grid transform/solve passes).
#include "pcnn_common.cuh"

namespace pcnn {

template <typename T> struct Cx;
template <> struct Cx<float> { using t = float2; };
template <> struct Cx<double> { using t = double2; };

template <typename C> __device__ __forceinline__ C cmk(decltype(C::x) a, decltype(C::x) b) { C r; r.x = a; r.y = b; return r; }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { return cmk<C>(a.x + b.x, a.y + b.y); }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { return cmk<C>(a.x - b.x, a.y - b.y); }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) { return cmk<C>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename C> __device__ __forceinline__ C cmul_mi(C a) { return cmk<C>(a.y, -a.x); }    // a * (-i)
template <typename C> __device__ __forceinline__ C cmul_pi(C a) { return cmk<C>(-a.y, a.x); }    // a * (+i)

__device__ __forceinline__ int spos(int idx) { return idx + (idx >> 3); }

// forward DFTs (sign -1), natural order in and out
template <typename C> __device__ __forceinline__ void dft2(C& a, C& b) { const C t = a; a = cadd(t, b); b = csub(t, b); }
template <typename C> __device__ __forceinline__ void dft4(C& v0, C& v1, C& v2, C& v3) {
    const C s02 = cadd(v0, v2), d02 = csub(v0, v2), s13 = cadd(v1, v3), d13 = csub(v1, v3);
    v0 = cadd(s02, s13);
    v2 = csub(s02, s13);
    v1 = cadd(d02, cmul_mi(d13));    // v0 - i v1 - v2 + i v3
    v3 = cadd(d02, cmul_pi(d13));    // v0 + i v1 - v2 - i v3
}
template <typename C> __device__ __forceinline__ void dft8(C (&u)[8]) {
    using R = decltype(C::x);
    C e0 = u[0], e1 = u[2], e2 = u[4], e3 = u[6], o0 = u[1], o1 = u[3], o2 = u[5], o3 = u[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    const R h = (R)0.70710678118654752440;
    o1 = cmk<C>((o1.x + o1.y) * h, (o1.y - o1.x) * h);        // * (1 - i)/sqrt2
    o2 = cmul_mi(o2);                                         // * (-i)
    o3 = cmk<C>((o3.y - o3.x) * h, -(o3.x + o3.y) * h);       // * (-1 - i)/sqrt2
    u[0] = cadd(e0, o0); u[4] = csub(e0, o0);
    u[1] = cadd(e1, o1); u[5] = csub(e1, o1);
    u[2] = cadd(e2, o2); u[6] = csub(e2, o2);
    u[3] = cadd(e3, o3); u[7] = csub(e3, o3);
}
template <typename C, int R> __device__ __forceinline__ void dftR(C (&u)[R]) {
    if constexpr (R == 8) dft8(u);
    else if constexpr (R == 4) dft4(u[0], u[1], u[2], u[3]);
    else dft2(u[0], u[1]);
}

template <typename C> __device__ __forceinline__ C ldgc(const C* p) { return *p; }
template <> __device__ __forceinline__ float2 ldgc<float2>(const float2* p) { return __ldg(p); }
template <> __device__ __forceinline__ double2 ldgc<double2>(const double2* p) { return __ldg(p); }

// One Stockham stage of radix R = 2^LR on the transform that lives at s (padded indexing); tid in [0, L/8).
// lp = log2 of the product of the radices already applied.  Ends with a CTA barrier.
template <typename C, int LR>
__device__ __forceinline__ void fft_stage(C* s, int tid, int q, int lp, const C* __restrict__ tw) {
    constexpr int R = 1 << LR, NB = 8 / R;
    const int NT = 1 << (q - 3), Tn = 1 << (q - LR), p = 1 << lp;
    C v[8];
    int jout[NB];
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int i = tid + m * NT;
        const int k = i & (p - 1);
        jout[m] = ((i - k) << LR) + k;
        C u[R];
#pragma unroll
        for (int t = 0; t < R; ++t) u[t] = s[spos(i + t * Tn)];
        if (lp > 0) {
            // w = exp(-2 pi i k / (p R)) from the per-stage table (entry [0][k] at offset p - 8: lanes with consecutive k read
            // consecutive entries), its powers by multiplication: one table load per butterfly instead of R-1.  With all
            // R-1 loaded, twiddles + chirp + spectrum tables (160 KB at L = 4096, double) overflowed what two resident
            // CTAs leave of the L1 and the kernel sat in long-scoreboard stalls (ncu: profiles/r02_checks_full_raw.csv).
            const C w1 = ldgc(tw + (p - 8) + k);
            u[1] = cmul(u[1], w1);
            if constexpr (R >= 4) {
                const C w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                u[2] = cmul(u[2], w2);
                u[3] = cmul(u[3], w3);
                if constexpr (R == 8) {
                    const C w4 = cmul(w2, w2), w5 = cmul(w4, w1), w6 = cmul(w3, w3), w7 = cmul(w4, w3);
                    u[4] = cmul(u[4], w4);
                    u[5] = cmul(u[5], w5);
                    u[6] = cmul(u[6], w6);
                    u[7] = cmul(u[7], w7);
                }
            }
        }
        dftR<C, R>(u);
#pragma unroll
        for (int t = 0; t < R; ++t) v[m * R + t] = u[t];
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NB; ++m)
#pragma unroll
        for (int t = 0; t < R; ++t) s[spos(jout[m] + (t << lp))] = v[m * R + t];
    __syncthreads();
}

// forward FFT of length 2^q in place in shared memory (data must be visible: barrier before the call)
template <typename C>
__device__ __forceinline__ void fft_fwd(C* s, int tid, int q, const C* __restrict__ tw) {
    const int a = q / 3, r = q - 3 * a;
    int lp = 0;
    for (int st = 0; st < a; ++st, lp += 3) fft_stage<C, 3>(s, tid, q, lp, tw);
    if (r == 1) fft_stage<C, 1>(s, tid, q, lp, tw);
    else if (r == 2) fft_stage<C, 2>(s, tid, q, lp, tw);
}

template <typename T>
struct DstPlan {                  // device tables of one transform length n (built by dst_plan_kernel)
    const typename Cx<T>::t* tw;  // [L]   per-stage twiddle tables (dst_twiddle_kernel)
    const typename Cx<T>::t* bh;  // [L]   FFT(conj chirp, wrapped) / L
    const typename Cx<T>::t* ch;  // [n+1] chirp c_m = exp(i pi m^2 / (2(n+1)))
    const double* lam;            // [n]   2 - 2 cos(k pi / (n+1)), k = 1..n
    int n, q;                     // L = 1 << q
};

// With a_idx = x_idx * c_idx stored at s[spos(idx)] for idx in [0, L) (zero outside 1..n) and a barrier behind it, leaves
// r = FFT(conj(FFT(a) * bh)) in s; S_k = Im(c_k * conj(r_k)) = c_k.y * r_k.x - c_k.x * r_k.y.
template <typename T>
__device__ __forceinline__ void dst_core(typename Cx<T>::t* s, int tid, const DstPlan<T>& pl) {
    using C = typename Cx<T>::t;
    const int NT = 1 << (pl.q - 3);
    fft_fwd<C>(s, tid, pl.q, pl.tw);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int idx = tid + m * NT;
        const C p = cmul(s[spos(idx)], ldgc(pl.bh + idx));
        s[spos(idx)] = cmk<C>(p.x, -p.y);
    }
    __syncthreads();
    fft_fwd<C>(s, tid, pl.q, pl.tw);
}

template <typename T>
__device__ __forceinline__ T dst_value(const typename Cx<T>::t* s, int k, const DstPlan<T>& pl) {
    const typename Cx<T>::t r = s[spos(k)], c = ldgc(pl.ch + k);
    return c.y * r.x - c.x * r.y;
}

// ---- plan construction: one CTA of L/8 threads, dynamic shared memory 9/8 L complex (tw was written by the launch before)
template <typename T>
__global__ void dst_plan_kernel(const typename Cx<T>::t* __restrict__ tw, typename Cx<T>::t* bh, typename Cx<T>::t* ch,
                                double* lam, int n, int q) {
    using C = typename Cx<T>::t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* s = reinterpret_cast<C*>(smem_raw);
    const int L = 1 << q, tid = threadIdx.x, NT = blockDim.x;
    const long long M = n + 1;
    for (int m = tid; m <= n; m += NT) {
        const long long r = ((long long)m * m) % (4 * M);            // exact argument reduction: period of m^2/(2M) is 4M
        double sn, cs;
        sincospi((double)r / (double)(2 * M), &sn, &cs);
        ch[m] = cmk<C>((T)cs, (T)sn);
        if (m >= 1) lam[m - 1] = 2.0 - 2.0 * cospi((double)m / (double)M);
    }
    for (int idx = tid; idx < L; idx += NT) {
        const int m = (idx < n) ? idx : ((L - idx < n) ? L - idx : -1);   // b_m = conj(c_|m|), |m| <= n-1, wrapped mod L
        C v = cmk<C>((T)0, (T)0);
        if (m >= 0) {
            const long long r = ((long long)m * m) % (4 * M);
            double sn, cs;
            sincospi((double)r / (double)(2 * M), &sn, &cs);
            v = cmk<C>((T)cs, (T)(-sn));
        }
        s[spos(idx)] = v;
    }
    __syncthreads();
    fft_fwd<C>(s, tid, q, tw);
    const T inv = (T)(1.0 / (double)L);
    for (int idx = tid; idx < L; idx += NT) {
        const C v = s[spos(idx)];
        bh[idx] = cmk<C>(v.x * inv, v.y * inv);
    }
}

// The twiddle tables, in their own launch (the FFT of the plan kernel reads them through the read-only data path, which is
// only coherent across kernel boundaries).  Stage with p = 2^lp sub-transforms already merged (lp = 3, 6, ...) and radix R
// (8, or 2 / 4 for the last stage) owns entries [p - 8, p - 8 + (R-1) p): tw[p - 8 + (t-1) p + k] = exp(-2 pi i t k / (p R)).
// Total <= L entries.
template <typename T>
__global__ void dst_twiddle_kernel(typename Cx<T>::t* tw, int q) {
    using C = typename Cx<T>::t;
    const int a = q / 3, r = q - 3 * a;
    for (int lp = 3; lp <= 3 * a; lp += 3) {
        const int LR = (lp < 3 * a) ? 3 : r;
        if (LR == 0) break;
        const int p = 1 << lp, R = 1 << LR;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < (R - 1) * p; e += gridDim.x * blockDim.x) {
            const int t = e / p + 1, k = e - (t - 1) * p;
            double sn, cs;
            sincospi(-2.0 * (double)((long long)t * k) / (double)((long long)p * R), &sn, &cs);
            tw[p - 8 + e] = cmk<C>((T)cs, (T)sn);
        }
    }
}

// ---- P1 / P3: lines along y (contiguous).  F = blockDim / (L/8) lines per CTA.
// MODE 0: x_j = -dx^2 f + adjacent Dirichlet values -> T[b][i-1][k-1] = S_k
// MODE 1: x_j = T[b][i-1][j-1] -> out[b][i][k] = S_k (fp32) and the Dirichlet ring (write order of multigrid.py:145-148:
//         top/bottom first, then left/right, i.e. the corners hold left/right values)
template <typename T, int MODE>
__global__ void __launch_bounds__(512, 2) dst_rows_kernel(const float* __restrict__ rhs, const float* __restrict__ left,
                                                       const float* __restrict__ top, const float* __restrict__ right,
                                                       const float* __restrict__ bottom, const float* __restrict__ dx,
                                                       double* __restrict__ tbuf, float* __restrict__ out, int B, int nx, int ny,
                                                       DstPlan<T> pl) {
    using C = typename Cx<T>::t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int q = pl.q, L = 1 << q, NT = L >> 3, Lp = L + (L >> 3);
    const int f = threadIdx.x >> (q - 3), tid = threadIdx.x & (NT - 1), F = blockDim.x >> (q - 3);
    C* s = reinterpret_cast<C*>(smem_raw) + (size_t)f * Lp;
    const int mx = nx - 2, my = ny - 2;
    const long long line = (long long)blockIdx.x * F + f, nlines = (long long)B * mx;
    const bool live = line < nlines;
    const int b = live ? (int)(line / mx) : 0, i = live ? (int)(line - (long long)b * mx) + 1 : 1;
    T h2 = (T)0;
    if (MODE == 0) { const T h = (T)dx[b]; h2 = -h * h; }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int j = tid + m * NT;
        C a = cmk<C>((T)0, (T)0);
        if (live && j >= 1 && j <= my) {
            T v;
            if (MODE == 0) {
                v = h2 * (T)__ldg(rhs + ((long long)b * nx + i) * ny + j);
                if (j == 1) v += (T)bottom[(long long)b * nx + i];
                if (j == my) v += (T)top[(long long)b * nx + i];
                if (i == 1) v += (T)left[(long long)b * ny + j];
                if (i == mx) v += (T)right[(long long)b * ny + j];
            } else {
                v = (T)tbuf[((long long)b * mx + (i - 1)) * my + (j - 1)];
            }
            const C c = ldgc(pl.ch + j);
            a = cmk<C>(v * c.x, v * c.y);
        }
        s[spos(j)] = a;
    }
    __syncthreads();
    dst_core<T>(s, tid, pl);
    if (!live) return;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int k = tid + m * NT;
        if (k >= 1 && k <= my) {
            const T v = dst_value<T>(s, k, pl);
            if (MODE == 0) tbuf[((long long)b * mx + (i - 1)) * my + (k - 1)] = (double)v;
            else out[((long long)b * nx + i) * ny + k] = (float)v;
        }
    }
    if (MODE == 1) {
        float* o = out + (long long)b * nx * ny;
        if (tid == 0) {
            o[(long long)i * ny] = bottom[(long long)b * nx + i];
            o[(long long)i * ny + ny - 1] = top[(long long)b * nx + i];
        }
        if (i == 1) for (int j = tid; j < ny; j += NT) o[j] = left[(long long)b * ny + j];
        if (i == mx) for (int j = tid; j < ny; j += NT) o[(long long)(nx - 1) * ny + j] = right[(long long)b * ny + j];
    }
}

// ---- P2: tridiagonal solves along x, one thread per (sample, y-mode k); consecutive threads own consecutive k, so every
// row access is one coalesced segment.  System per column: -u[i-1] + d u[i] - u[i+1] = r[i], d = 2 + lam_y[k]
// (from (lam_x + lam_y) u^ = b^ with the x-direction left in physical space).  Thomas:
//   forward   inv_i = 1 / (d + c_{i-1}),  c_i = -inv_i,  r'_i = (r_i + r'_{i-1}) inv_i          (c_0 = r'_0 = 0)
//   backward  u_n = r'_n,  u_i = r'_i - c_i u_{i+1}
// The elimination coefficients c_i depend on (i, k) only: the threads of sample 0 store them once per call ([mx][my]
// doubles, L2-resident for the backward sweep of every sample).  `scale` = the inverse-DST normalisation 2/(my+1) of P3.
constexpr int kThomasUnroll = 16;      // independent row loads in flight per thread

__global__ void __launch_bounds__(128) dst_thomas_forward_kernel(double* __restrict__ tbuf, double* __restrict__ ctab,
                                                                const double* __restrict__ lamy, int mx, int my) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= my) return;
    const int b = blockIdx.y;
    double* col = tbuf + (long long)b * mx * my + k;
    const double d = 2.0 + lamy[k];
    double c = 0.0, rp = 0.0;
    for (int i0 = 0; i0 < mx; i0 += kThomasUnroll) {
        double v[kThomasUnroll];
#pragma unroll
        for (int u = 0; u < kThomasUnroll; ++u) v[u] = (i0 + u < mx) ? col[(long long)(i0 + u) * my] : 0.0;
#pragma unroll
        for (int u = 0; u < kThomasUnroll; ++u) {
            if (i0 + u < mx) {
                const double inv = 1.0 / (d + c);
                rp = (v[u] + rp) * inv;
                c = -inv;
                col[(long long)(i0 + u) * my] = rp;
                if (b == 0) ctab[(long long)(i0 + u) * my + k] = c;
            }
        }
    }
}

__global__ void __launch_bounds__(128) dst_thomas_backward_kernel(double* __restrict__ tbuf, const double* __restrict__ ctab,
                                                                 int mx, int my, double scale) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= my) return;
    double* col = tbuf + (long long)blockIdx.y * mx * my + k;
    const double* cc = ctab + k;
    double un = 0.0;                    // u_{i+1}; the last row has no successor (c_n multiplies nothing)
    for (int i0 = mx - 1; i0 >= 0; i0 -= kThomasUnroll) {
        double v[kThomasUnroll], cv[kThomasUnroll];
#pragma unroll
        for (int u = 0; u < kThomasUnroll; ++u) {
            const bool ok = i0 - u >= 0;
            v[u] = ok ? col[(long long)(i0 - u) * my] : 0.0;
            cv[u] = ok ? __ldg(cc + (long long)(i0 - u) * my) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kThomasUnroll; ++u) {
            if (i0 - u >= 0) {
                un = (i0 - u == mx - 1) ? v[u] : v[u] - cv[u] * un;
                col[(long long)(i0 - u) * my] = un * scale;
            }
        }
    }
}

static int dst_log2_len(int n) {       // q with 2^q >= max(8, 2n-1)
    int q = 3;
    while ((1 << q) < 2 * n - 1) ++q;
    return q;
}

template <typename T>
static size_t plan_bytes(int n) {
    const int L = 1 << dst_log2_len(n);
    size_t b = (size_t)(2 * L + n + 1) * 2 * sizeof(T);
    b = (b + 15) / 16 * 16;
    return b + (size_t)n * sizeof(double);
}

template <typename T>
static DstPlan<T> plan_view(void* plan, int n) {
    using C = typename Cx<T>::t;
    DstPlan<T> p;
    const int q = dst_log2_len(n), L = 1 << q;
    C* base = reinterpret_cast<C*>(plan);
    p.tw = base;
    p.bh = base + L;
    p.ch = base + 2 * L;
    size_t b = (size_t)(2 * L + n + 1) * 2 * sizeof(T);
    b = (b + 15) / 16 * 16;
    p.lam = reinterpret_cast<const double*>(reinterpret_cast<unsigned char*>(plan) + b);
    p.n = n;
    p.q = q;
    return p;
}

template <typename T>
static int plan_init(void* plan, int n, cudaStream_t st) {
    using C = typename Cx<T>::t;
    DstPlan<T> p = plan_view<T>(plan, n);
    const int L = 1 << p.q, NT = L / 8;
    const size_t smem = (size_t)(L + L / 8) * sizeof(C);
    dst_twiddle_kernel<T><<<ceil_div(L, 256), 256, 0, st>>>(const_cast<C*>(p.tw), p.q);
    PCNN_CHECK_LAUNCH();
    PCNN_CHECK_CUDA(cudaFuncSetAttribute(dst_plan_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dst_plan_kernel<T><<<1, NT, smem, st>>>(p.tw, const_cast<C*>(p.bh), const_cast<C*>(p.ch),
                                            const_cast<double*>(p.lam), n, p.q);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

template <typename T>
static int solve(const float* rhs, const float* left, const float* top, const float* right, const float* bottom,
                 const float* dx, void* plan_y, void* work, float* out, int B, int nx, int ny, cudaStream_t st) {
    using C = typename Cx<T>::t;
    const int mx = nx - 2, my = ny - 2;
    DstPlan<T> py = plan_view<T>(plan_y, my);
    double* tbuf = reinterpret_cast<double*>(work);
    double* ctab = tbuf + (size_t)B * mx * my;
    const int NT = 1 << (py.q - 3);
    const int threads = NT > 256 ? NT : 256, F = threads / NT;
    const size_t smem = (size_t)threads * 9 * sizeof(C);
    const long long nlines = (long long)B * mx;
    const unsigned grid = (unsigned)((nlines + F - 1) / F);
    PCNN_CHECK_CUDA(cudaFuncSetAttribute(dst_rows_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PCNN_CHECK_CUDA(cudaFuncSetAttribute(dst_rows_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dst_rows_kernel<T, 0><<<grid, threads, smem, st>>>(rhs, left, top, right, bottom, dx, tbuf, out, B, nx, ny, py);
    PCNN_CHECK_LAUNCH();
    const dim3 tg((unsigned)ceil_div(my, 128), (unsigned)B);
    dst_thomas_forward_kernel<<<tg, 128, 0, st>>>(tbuf, ctab, py.lam, mx, my);
    PCNN_CHECK_LAUNCH();
    dst_thomas_backward_kernel<<<tg, 128, 0, st>>>(tbuf, ctab, mx, my, 2.0 / (my + 1));
    PCNN_CHECK_LAUNCH();
    dst_rows_kernel<T, 1><<<grid, threads, smem, st>>>(rhs, left, top, right, bottom, dx, tbuf, out, B, nx, ny, py);
    PCNN_CHECK_LAUNCH();
    return PCNN_OK;
}

}  // namespace pcnn

using namespace pcnn;

extern "C" size_t pcnn_dst_fft_plan_bytes(int n, int use_double) {
    if (n < 1 || n > 2048) return 0;
    return use_double ? plan_bytes<double>(n) : plan_bytes<float>(n);
}

extern "C" int pcnn_dst_fft_plan_init(void* plan, int n, int use_double, void* stream) {
    PCNN_CHECK_ARG(plan && n >= 1 && n <= 2048, "dst_fft_plan_init: needs 1 <= n <= 2048 interior points per side");
    return use_double ? plan_init<double>(plan, n, (cudaStream_t)stream) : plan_init<float>(plan, n, (cudaStream_t)stream);
}

extern "C" size_t pcnn_dst_fft_workspace_bytes(int B, int nx, int ny) {
    if (B <= 0 || nx < 3 || ny < 3) return 0;
    return ((size_t)B + 1) * (nx - 2) * (ny - 2) * sizeof(double);      // T for every sample + one table of elimination coefficients
}

extern "C" int pcnn_dst_fft_passes(void) { return 3; }

extern "C" int pcnn_dst_solve_fft(const float* rhs, const float* left, const float* top, const float* right,
                                  const float* bottom, const float* dx, void* plan_y, void* work, float* out, int B,
                                  int nx, int ny, int use_double, void* stream) {
    PCNN_CHECK_ARG(rhs && left && top && right && bottom && dx && plan_y && work && out, "dst_solve_fft: null pointer");
    PCNN_CHECK_ARG(B > 0 && B <= 65535 && nx >= 3 && ny >= 3 && ny <= 2050, "dst_solve_fft: needs nx >= 3, 3 <= ny <= 2050 and B <= 65535");
    if (use_double)
        return solve<double>(rhs, left, top, right, bottom, dx, plan_y, work, out, B, nx, ny, (cudaStream_t)stream);
    return solve<float>(rhs, left, top, right, bottom, dx, plan_y, work, out, B, nx, ny, (cudaStream_t)stream);
}
