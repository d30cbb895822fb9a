/* pcnn.h -- C ABI of libpcnn.so: hand-written sm_100a CUDA kernels for the batched-inference hot
 * path of aligirayhanozbay/poisson_CNN (Poisson_CNN_Legacy forward, FD Laplacian residual, DST
 * direct solve).
 *
 * The reference is pure Python on TensorFlow and has no FFI of its own; each entry point below
 * replaces the TensorFlow op call site(s) cited next to it (paths relative to
 * /root/reference/poisson_CNN/).  The host side (poisson_cnn_b200/, Python, ctypes) mirrors the
 * reference's poisson_CNN.models call surface and drives these functions.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the library never allocates device memory and never synchronises: all work is enqueued on
 *     the caller's stream (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - the caller owns every buffer; the library borrows them for the duration of the call;
 *   - every function returns 0 on success or a negative pcnn_status; pcnn_last_error() returns a
 *     thread-local message for the last failure.  No C++ exception crosses the ABI;
 *   - tensors are NCHW ("channels_first", the reference's data_format), fp32, dense unless a
 *     *_bstride (batch stride, in elements) argument says otherwise.
 */
#ifndef PCNN_H_
#define PCNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCNN_VERSION 200

typedef enum {
    PCNN_OK = 0,
    PCNN_ERR_INVALID_ARGUMENT = -1,
    PCNN_ERR_CUDA = -2,
    PCNN_ERR_UNSUPPORTED = -3
} pcnn_status;

/* activations: utils/convert_tf_object_names.py evals "tf.nn.leaky_relu" (alpha 0.2) / "tf.nn.tanh" */
enum { PCNN_ACT_LINEAR = 0, PCNN_ACT_LEAKY_RELU = 1, PCNN_ACT_TANH = 2 };
/* tf.pad modes: utils/apply_advanced_padding_and_call_conv_layer.py:6,18 */
enum { PCNN_PAD_CONSTANT = 0, PCNN_PAD_SYMMETRIC = 1, PCNN_PAD_REFLECT = 2 };
/* SpatialPyramidPool pooling_type: layers/SpatialPyramidPool.py:11-14 */
enum { PCNN_POOL_AVG = 0, PCNN_POOL_MAX = 1 };
/* HPNN bc_type: models/Homogeneous_Poisson_NN_Legacy.py:106-113 */
enum { PCNN_BC_DIRICHLET = 0, PCNN_BC_NEUMANN = 1 };

int pcnn_version(void);
const char* pcnn_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
long long pcnn_launch_count(void);

/* ---- convolution ---------------------------------------------------------------------------
 * tf.pad(mode) + Conv2D(VALID) + bias + activation, then (optionally, in this order) the
 * inference BatchNorm affine that follows the conv, the residual add of blocks/resnet.py:37, and a
 * per-(sample,channel) scale (the dx-MLP einsum, models/Homogeneous_Poisson_NN_Legacy.py:231).
 * Replaces utils/apply_advanced_padding_and_call_conv_layer.py:17-20, blocks/resnet.py:29-39 and
 * the Keras BatchNormalization calls.  Conv1D (models/Dirichlet_BC_NN_Legacy.py:52-64) is the
 * H == 1, kh == 1 case.
 *   in        [B, Cin, H, W]   (batch stride in_bstride elements)
 *   kernel    [kh, kw, Cin, Cout]  Keras layout; cross-correlation, pad left k/2, right k/2-(1-k%2)
 *   bias      [Cout] or NULL
 *   bn_scale, bn_shift [Cout] or NULL:  v = v*bn_scale + bn_shift  (gamma/sqrt(var+1e-3), beta-mean*scale)
 *   residual  [B, Cout, H, W] or NULL (batch stride res_bstride), added after BN
 *   out_scale [B, Cout] or NULL, multiplied last
 *   out       [B, Cout, H, W]  (batch stride out_bstride)
 */
int pcnn_conv2d_f32(const float* in, const float* kernel, const float* bias, const float* bn_scale,
                    const float* bn_shift, const float* residual, const float* out_scale, float* out,
                    int B, int Cin, int Cout, int H, int W, int kh, int kw, int pad_mode,
                    float pad_value, int act, int64_t in_bstride, int64_t out_bstride,
                    int64_t res_bstride, void* stream);

/* ---- pooling / upsampling -------------------------------------------------------------------
 * AveragePooling2D(pool=s, stride=s, 'same'): out = ceil(N/s), pad_before = floor((out*s-N)/2),
 * divisor = number of valid cells.  Replaces utils/get_pooling_method.py:3-6 as used by
 * blocks/bottleneck_block.py:36-37,72 and layers/Scaling.py:29.   out [B,C,ceil(H/s),ceil(W/s)] dense. */
int pcnn_avgpool_same_f32(const float* in, float* out, int B, int C, int H, int W, int s,
                          int64_t in_bstride, void* stream);

/* tf.nn.conv2d_transpose(in, kernel[kh,kw,Cout,Cin], out_shape, strides=stride, 'SAME') + bias +
 * activation (layers/deconvupscale.py:100-109).  out = (accumulate ? out : 0) + alpha * result, so
 * the 8-way merge sum of models/Homogeneous_Poisson_NN_Legacy.py:222 needs no extra pass. */
int pcnn_deconv_same_f32(const float* in, const float* kernel, const float* bias, float* out, int B,
                         int Cin, int Cout, int ih, int iw, int oh, int ow, int kh, int kw,
                         int stride, int act, float alpha, int accumulate, int64_t out_bstride,
                         void* stream);

/* tf.image.resize(method, antialias=False) (layers/Upsample.py:56-59), separable gather form:
 * out[y,x] = sum_{a,b} wy[y,a]*wx[x,b]*in[iy[y,a], ix[x,b]].  The per-axis tables (taps = 1 nearest,
 * 2 bilinear, 4 bicubic incl. TF's 1024-step Keys table) are built by the host and live on the
 * device.  out = (accumulate ? out : 0) + alpha * result. */
int pcnn_resize_f32(const float* in, const int32_t* iy, const float* wy, const int32_t* ix,
                    const float* wx, int taps, float* out, int B, int C, int ih, int iw, int oh,
                    int ow, float alpha, int accumulate, int64_t out_bstride, void* stream);

/* SpatialPyramidPool.call (layers/SpatialPyramidPool.py:48-66): every bin is reduced over channels
 * AND space.  boxes [nbins,4] = (y0,y1,x0,x1) from split_indices (dataset/utils/split_indices.py);
 * out [B,nbins].  An empty box gives -inf (max) / NaN (avg) like tf.reduce_*. */
int pcnn_spp_f32(const float* in, const int32_t* boxes, float* out, int B, int C, int H, int W,
                 int nbins, int mode, void* stream);

/* tf.keras.layers.Dense: y = act(x @ kernel + bias); x [B,nin], kernel [nin,nout]. */
int pcnn_dense_f32(const float* x, const float* kernel, const float* bias, float* y, int B, int nin,
                   int nout, int act, void* stream);

/* ---- per-sample normalisation (dataset/utils/set_max_magnitude.py:4-50) ------------------- */
/* maxabs[b] = max |x[b,:]| ; n elements per sample */
int pcnn_maxabs_f32(const float* x, float* maxabs, int B, int64_t n, void* stream);
/* y[b,:] = x[b,:] * (1 / maxabs[b])   (the reference multiplies by the quotient 1.0/max) */
int pcnn_scale_inv_f32(const float* x, const float* maxabs, float* y, int B, int64_t n, void* stream);

/* ---- model glue ---------------------------------------------------------------------------- */
/* concat(rhs, cos(pi*linspace) along x, along y): models/Homogeneous_Poisson_NN_Legacy.py:172-180,196-198.
 * posx [H], posy [W] host-built tables; out [B,3,H,W]. */
int pcnn_hpnn_input_f32(const float* rhs, const float* posx, const float* posy, float* out, int B,
                        int H, int W, void* stream);
/* concat(bc, pos_nd[...,0,:]) -> [B,3,n] (models/Dirichlet_BC_NN_Legacy.py:132,136); posx0 = posx[0] */
int pcnn_dbcnn_input_f32(const float* bc, float posx0, const float* posy, float* out, int B, int n,
                         void* stream);
/* einsum('bmy,mx,bm->bmxy') + concat(pos_nd) (models/Dirichlet_BC_NN_Legacy.py:151-154):
 * h [B,M,n], sinh_basis [M,xres], modew [B,M] -> out [B,M+2,xres,n] */
int pcnn_dbcnn_expand_f32(const float* h, const float* sinh_basis, const float* modew,
                          const float* posx, const float* posy, float* out, int B, int M, int xres,
                          int n, void* stream);
/* out = raw * (1/maxabs[b]); out[:, :, 0, :] = bc   (models/Dirichlet_BC_NN_Legacy.py:158-160) */
int pcnn_dbcnn_finalize_f32(const float* raw, const float* maxabs, const float* bc, float* out, int B,
                            int xres, int n, void* stream);
/* Scaling multiply + boundary ring (layers/Scaling.py:55, models/Homogeneous_Poisson_NN_Legacy.py:251):
 * out = pad((y*(1+s[b]))[1:-1,1:-1], 1, CONSTANT 0 | SYMMETRIC); s may be NULL. */
int pcnn_hpnn_finalize_f32(const float* y, const float* s, float* out, int B, int H, int W,
                           int bc_type, int64_t y_bstride, void* stream);
/* Inputs of the two MLPs: out [B,3+nextra] = [dx, L0, L1, extra...] with L = dx*(n-1)
 * (models/Homogeneous_Poisson_NN_Legacy.py:193,202); normalize != 0 divides L0,L1 by max(L0,L1)
 * (models/Dirichlet_BC_NN_Legacy.py:129-130,142; extra = the SPP vector). */
int pcnn_dense_input_f32(const float* dx, const float* extra, float* out, int B, int n0, int n1,
                         int nextra, int normalize, void* stream);
/* Poisson_CNN_Legacy.call tail (models/Poisson_CNN_Legacy.py:30-47) with the rot90/flip of
 * dataset/utils/flip_and_rotate_tensor.py folded into the indexing:
 * pred[b,i,j] = L[b,i,j]*ml + R[b,nx-1-i,j]*mr + T[b,ny-1-j,i]*mt + Bt[b,j,i]*mb
 *             + hp[b,i,j] * (dx[b]*(max(nx,ny)-1))^2 * mrhs
 * L,R [B,nx,ny]; T,Bt [B,ny,nx]; m* [B] = max|.| of the raw inputs (1/scaling factor). */
int pcnn_merge_f32(const float* hp, const float* L, const float* T, const float* R, const float* Bt,
                   const float* dx, const float* mrhs, const float* ml, const float* mt,
                   const float* mr, const float* mb, float* out, int B, int nx, int ny, void* stream);

/* ---- checks -------------------------------------------------------------------------------- */
/* linear_operator_loss.__call__ (losses/physics_informed_loss.py:35-50): central-difference
 * Laplacian (stencil 3: [1,-2,1]; stencil 5: [-1/12,4/3,-5/2,4/3,-1/12] per direction, weights
 * 1/dx_d^2, dataset/utils/build_fd_coefficients.py:5-42) of sol against the interior of rhs.
 * grid_spacings [B,2].  sq_sum[b] = sum over interior of (rhs - Lap(sol))^2 (double),
 * divided by rhs_maxabs[b]^2 when rhs_maxabs != NULL (the reference's normalize=True).
 * The loss is sum(sq_sum)/(B*interior points). */
int pcnn_laplacian_residual_f32(const float* rhs, const float* sol, const float* grid_spacings,
                                const float* rhs_maxabs, double* sq_sum, int B, int H, int W,
                                int stencil, void* stream);
/* JacobiIterationLayer (layers/JacobiIterationLayer.py:44-66), stencil [3,3] orders [2,2]:
 * one sweep cur -> next; ring copied unchanged.  grid_spacings [B,2]. */
int pcnn_jacobi_sweep_f32(const float* cur, const float* rhs, const float* grid_spacings, float* next,
                          int B, int H, int W, void* stream);

/* Direct solve of the reference's ground-truth system (dataset/solvers/multigrid.py:98-150,
 * dataset/solvers/cholesky.py:45-119) with a DST-I eigen-decomposition, in double precision:
 * A u = -dx^2 f + adjacent BCs on the interior, ring := BCs (left/right written last).
 * rhs [B,nx,ny] fp32; left,right [B,ny]; top,bottom [B,nx]; dx [B]; out [B,nx,ny] fp32.
 * sx [(nx-2)^2], sy [(ny-2)^2] double sine matrices and work (2*B*(nx-2)*(ny-2) doubles) are
 * caller-provided (pcnn_dst_workspace_bytes / pcnn_dst_sine_matrix). */
size_t pcnn_dst_workspace_bytes(int B, int nx, int ny);
int pcnn_dst_sine_matrix(double* s, int m, void* stream);
int pcnn_dst_solve(const float* rhs, const float* left, const float* top, const float* right,
                   const float* bottom, const float* dx, const double* sx, const double* sy,
                   double* work, float* out, int B, int nx, int ny, void* stream);

/* The same solve in O(N^2 log N): DST-I along y for every row (Bluestein chirp-z transform on power-of-two FFTs in shared
 * memory, so ANY grid size works: N-1 = 255, 1023, 2047 are not FFT-friendly), then -- the system having decoupled per y-mode
 * into constant-coefficient tridiagonal systems along x -- one float64 Thomas solve per (sample, mode), then the inverse DST
 * along y (csrc/dst_fft.cu).  Three full-grid passes (4 launches) instead of four dense O(N^3) sine-matrix products.
 * A plan holds the device tables of the interior row length n = ny-2 (FFT twiddles, spectrum of the conjugate chirp, chirp,
 * eigenvalues): pcnn_dst_fft_plan_bytes bytes, filled once by pcnn_dst_fft_plan_init.  use_double selects the FFT arithmetic
 * (1 = float64 like the reference's solver, 0 = float32); the intermediate grid and the tridiagonal solves are always
 * float64.  work: pcnn_dst_fft_workspace_bytes.  nx >= 3, 3 <= ny <= 2050.  pcnn_dst_fft_passes() = full-grid passes made
 * (3), reported by bench.py next to the 8 B/point figure. */
size_t pcnn_dst_fft_plan_bytes(int n, int use_double);
int pcnn_dst_fft_plan_init(void* plan, int n, int use_double, void* stream);
size_t pcnn_dst_fft_workspace_bytes(int B, int nx, int ny);
int pcnn_dst_fft_passes(void);
int pcnn_dst_solve_fft(const float* rhs, const float* left, const float* top, const float* right,
                       const float* bottom, const float* dx, void* plan_y, void* work, float* out, int B,
                       int nx, int ny, int use_double, void* stream);

/* ---- the caller after the path: pressure-Poisson solve of the reference's Navier-Stokes projection solver -------------
 * (Navier_Stokes_2D/solvers.py:153-186 operator, :204-334 solve).  A = minus the cell-centred 5-point
 * Laplacian with homogeneous Neumann boundaries, (A v)[i,j] = (deg v[i,j] - sum of existing neighbours) / dx^2, on [B,H,W]
 * fields with per-sample dx [B].  pcnn_neumann_cg_solve runs batched conjugate gradients for  A p = -rhs - mean(-rhs)
 * (the reference's zero-integral Lagrange row with uniform Riemann weights) and returns the zero-mean p:
 *   x            in: initial guess (e.g. the Neumann HPNN's prediction), multiplied by guess_scale[b] when given (the
 *                reference's ((dx (n-1))^2 / sf) rescale of the network output, solvers.py:250); out: solution
 *   max_iter     CG iterations enqueued (2 kernels each; per-sample step lengths live on the device, a sample whose
 *                residual is below rel_tol * |b| freezes); no host synchronisation inside
 *   residual_history  [max_iter][B] doubles or NULL: |r_k| / |b| after every iteration
 *   workspace    pcnn_neumann_cg_workspace_bytes(B,H,W), 8-byte aligned */
size_t pcnn_neumann_cg_workspace_bytes(int B, int H, int W);
int pcnn_neumann_laplacian_apply_f32(const float* v, const float* dx, float* out, int B, int H, int W, void* stream);
int pcnn_neumann_cg_solve(const float* rhs, const float* dx, const float* guess_scale, float* x, int B, int H, int W,
                          int max_iter, double rel_tol, double* residual_history, void* workspace, void* stream);

/* Fused 1-D convolution stack (csrc/boundary_stack.cu): n_layers Conv1D layers (Keras kernels [k][Cin][Cout],
 * odd k <= 19, <= 28 channels) applied back to back to every signal of in [B][Cin0][n], all activations staying in
 * shared memory.  Per layer: tf.pad(pad_mode) -> conv -> bias -> act -> BN affine (bn_scale/bn_shift may be NULL per
 * layer); flags[l] & 1 saves the layer's INPUT, flags[l] & 2 adds the saved tensor after act/BN -- a resnet
 * (blocks/resnet.py:29-39) is the three layers {1, 2, 0}.  Replaces the boundary_convolution_ops loop of
 * models/Dirichlet_BC_NN_Legacy.py:136-141 (32 launches).  The pointer arrays are HOST arrays of device pointers. */
int pcnn_boundary_stack_f32(const float* in, float* out, int B, int n, int Cin0, int n_layers,
                            const float* const* kernels, const float* const* biases,
                            const float* const* bn_scale, const float* const* bn_shift, const int* ksize,
                            const int* cin, const int* cout, const int* flags, int act, int pad_mode,
                            float pad_value, void* stream);

/* The same fused layer program on tiny 2-D maps, H*W <= 64 (csrc/smallmap_stack.cu): the conv + resnet chain of a
 * bottleneck branch whose pooled map is 2x2 .. 8x8 (blocks/bottleneck_block.py:36-50 at downsampling 32/64/128).
 * Keras Conv2D kernels [k][k][Cin][Cout], odd k <= 7, <= 32 channels; in [B][Cin0][H][W] -> out [B][Cout_last][H][W]. */
int pcnn_smallmap_stack_f32(const float* in, float* out, int B, int H, int W, int Cin0, int n_layers,
                            const float* const* kernels, const float* const* biases,
                            const float* const* bn_scale, const float* const* bn_shift, const int* ksize,
                            const int* cin, const int* cout, const int* flags, int act, int pad_mode,
                            float pad_value, void* stream);

/* ---- tensor-core path (tcgen05 / TMEM / TMA bulk copies), csrc/conv_tc.cu ---------------------
 * Activations live in the "BLK8" layout: fp16 [B][Cpad/8][H+14][W+14][8] with Cpad = round_up(C,16)
 * and a 7-pixel halo materialised in memory (zero = CONSTANT padding; pcnn_blk8_halo_fill mirrors it
 * for SYMMETRIC).  pcnn_blk8_bytes gives the allocation size (incl. slack read by partial tiles);
 * buffers must start zeroed, the kernels never write the halo or the channel padding. */
size_t pcnn_blk8_bytes(int B, int C, int H, int W);
/* NCHW fp32 [B,C,H,W] (batch stride in_bstride) -> channels [c_offset, c_offset+C) of a BLK8 buffer
 * holding c_total channels (c_offset multiple of 8: this is how concat is assembled in place). */
/* mode = precision mode of the tensor (see pcnn_conv2d_tc): 1 fp16 only; 2 second buffer = fp16
 * remainder; 3 second buffer = e4m3 planes (2c: x, 2c+1: remainder * 2^11, 16 channels each).
 * Mode 3, tensors with an odd number of live 8-channel planes (c_total in 1..8, 17..24, ...): the lone last plane p
 * keeps [e4m3(x) x 8 | e4m3(remainder * 2^11) x 8] per pixel in q plane p and leaves q plane p+1 zero.  All
 * producers here (to_blk8, conv2d_tc, dbcnn_expand_blk8, upsample_merge_blk8) write and all consumers (conv2d_tc,
 * from_blk8) read that form; it belongs to the TENSOR (c_total), not to the channel range of one call. */
/* halo_mode: as out_halo_mode of pcnn_conv2d_tc (0 = halo untouched, SYMMETRIC = also write the mirrored ring). */
int pcnn_to_blk8(const float* in, void* out, void* out_lo, int mode, int B, int C, int H, int W, int c_total,
                 int c_offset, int64_t in_bstride, int halo_mode, void* stream);
int pcnn_from_blk8(const void* in, const void* in_lo, int mode, float* out, int B, int C, int H, int W,
                   int c_total, int c_offset, int64_t out_bstride, void* stream);
/* tf.pad ring of width pad (<= 7) around the interior: mode PCNN_PAD_CONSTANT writes zeros,
 * PCNN_PAD_SYMMETRIC mirrors (utils/apply_advanced_padding_and_call_conv_layer.py:18). */
int pcnn_blk8_halo_fill(void* buf, int B, int C, int H, int W, int pad, int mode, void* stream);
/* pcnn_dbcnn_expand_f32 writing the BLK8 layout directly (the [B,29,H,W] fp32 tensor never exists). */
int pcnn_dbcnn_expand_blk8(const float* h, const float* sinh_basis, const float* modew, const float* posx,
                           const float* posy, void* out, void* out_lo, int mode, int B, int M, int xres,
                           int n, void* stream);
/* Keras kernel [k,k,Cin,Cout] fp32 -> fp16 operand image of the row-group GEMM:
 * [ceil(Cin/16)][k][2][(k+2(RT-1))*CP][8] with CP = pcnn_conv_tc_channel_slots(Cout, k) channel slots per
 * output row (8, 16, 24 or 32) and RT = 128/CP (5 for CP = 24) output rows per tile (see conv_tc.cu).
 * When ceil(Cin/8) is odd (and k >= 3) the last chunk holds one live 8-channel plane and its stages pair two
 * column taps in the two K halves instead ((k+1)/2 stages used): 21 % fewer MMAs on 17..24-channel layers.
 * Done once per layer at load time. */
int pcnn_conv_tc_channel_slots(int Cout, int k);
size_t pcnn_conv_tc_packed_weight_bytes(int kh, int kw, int Cin, int Cout, int nsplit);
/* scale: a power of two applied to the weights before rounding (keeps W_hi and W_lo in fp16's
 * normal range); pass its reciprocal as acc_scale to pcnn_conv2d_tc, which undoes it exactly. */
int pcnn_conv_tc_pack_weights(const float* kernel, void* packed, int kh, int kw, int Cin, int Cout,
                              int nsplit, float scale, void* stream);
/* Fused upsample + merge of the HPNN bottleneck branches straight into a BLK8 tensor (csrc/upsample_merge.cu):
 *   out[:, c_offset : c_offset+C] = alpha * ( sum_d act_d(conv2d_transpose_d(in_d) + bias_d) + sum_r resize_r(in_r) )
 * i.e. blocks/bottleneck_block.py:57-118 (deconvupscale with k == stride, 'SAME'; Upsample via the per-axis
 * tables of pcnn_resize_f32) followed by the branch sum of models/Homogeneous_Poisson_NN_Legacy.py:226-233.
 * The pointer arrays are HOST arrays of n_deconv / n_resize device pointers (<= 8 each); deconv inputs are
 * [B,C,ih,iw] fp32 with PACKED kernels (pcnn_upsample_merge_pack_kernel); resize sources are [B,C,ih,iw] fp32,
 * C*ih*iw <= 8192, sources and tables 16-byte aligned.
 * C % 8 == 0, C <= 32; c_offset % 16 == 0; mode = precision mode of the destination (1, 2, 3). */
/* dc_kernel[d] of pcnn_upsample_merge_blk8: the Keras deconv kernel [s,s,C,C] re-laid once per layer as
 * [s][s][C/8][8*C+4] floats (phase, 8-channel group) blocks, so one bulk copy stages a row phase. */
size_t pcnn_upsample_merge_packed_floats(int stride, int C);
int pcnn_upsample_merge_pack_kernel(const float* kernel, float* packed, int stride, int C, void* stream);
int pcnn_upsample_merge_blk8(int n_deconv, const float* const* dc_in, const float* const* dc_kernel,
                             const float* const* dc_bias, const int* dc_stride, const int* dc_ih,
                             const int* dc_iw, const int* dc_act, int n_resize, const float* const* rs_in,
                             const int32_t* const* rs_iy, const float* const* rs_wy,
                             const int32_t* const* rs_ix, const float* const* rs_wx, const int* rs_taps,
                             const int* rs_ih, const int* rs_iw, float alpha, void* out, void* out_lo,
                             int mode, int B, int C, int H, int W, int c_total, int c_offset, void* stream);
/* pcnn_upsample_merge_blk8 with the transpose convolutions on the tensor cores (csrc/upsample_merge_tc.cu; mma.sync
 * m16n8k16, fp16 operands, fp32 accumulation): dc_in[d] are BLK8 fp16 tensors [B][4][ih+14][iw+14][8] (C = 32 channels, the
 * hi buffer of a branch output: no fp32 copy of the branch is made), dc_wpack[d] the Keras deconv kernel [s,s,32,32]
 * re-laid once per layer as fp16 [s][s][32][36] (pcnn_upsample_merge_tc_pack_kernel); resize branches as in
 * pcnn_upsample_merge_blk8.  Same destination conventions (mode 3 needs an even number of 8-channel planes). */
/* shared memory the tensor-core kernel needs for these branches (0: unsupported); beyond 227 KB use pcnn_upsample_merge_blk8 */
size_t pcnn_upsample_merge_tc_smem_bytes(int n_deconv, const int* dc_stride, int n_resize, const int* rs_ih, const int* rs_iw);
size_t pcnn_upsample_merge_tc_packed_bytes(int stride);
int pcnn_upsample_merge_tc_pack_kernel(const float* kernel, void* packed, int stride, void* stream);
int pcnn_upsample_merge_tc_blk8(int n_deconv, const void* const* dc_in, const void* const* dc_wpack,
                                const float* const* dc_bias, const int* dc_stride, const int* dc_ih,
                                const int* dc_iw, const int* dc_act, int n_resize, const float* const* rs_in,
                                const int32_t* const* rs_iy, const float* const* rs_wy,
                                const int32_t* const* rs_ix, const float* const* rs_wx, const int* rs_taps,
                                const int* rs_ih, const int* rs_iw, float alpha, void* out, void* out_lo,
                                int mode, int B, int H, int W, int c_total, int c_offset, void* stream);
/* In place: channels [c_offset, c_offset+C) of the BLK8 tensor `buf` += alpha * sum_r resize_r(rs_in[r]) -- the Upsample branches
 * of the merge (layers/Upsample.py:56-59, models/Homogeneous_Poisson_NN_Legacy.py:226-233) as a second pass after
 * pcnn_upsample_merge_tc_blk8 ran with the transpose-conv branches only: used when the low-resolution resize sources are too large
 * for that kernel's shared-memory staging (grids beyond ~400 pixels a side; before, such grids fell back to eight fp32
 * read-modify-write passes: 12.8 of 51.7 ms of a 2048^2 forward).  Branch arguments as in pcnn_upsample_merge_blk8; C % 8 == 0
 * (mode 3: C % 16 == 0 and no tail plane in the destination). */
int pcnn_resize_add_blk8(int n_resize, const float* const* rs_in, const int32_t* const* rs_iy, const float* const* rs_wy,
                         const int32_t* const* rs_ix, const float* const* rs_wx, const int* rs_taps, const int* rs_ih,
                         const int* rs_iw, float alpha, void* buf, void* buf_lo, int mode, int B, int C, int H, int W,
                         int c_total, int c_offset, void* stream);
/* Same operator as pcnn_conv2d_f32 (pad + VALID conv + bias + act [+BN] [+residual] [*out_scale]) on
 * tcgen05 tensor cores: FP16 operands, FP32 accumulation in TMEM.  in/out/residual are BLK8 buffers
 * with Cin_total / Cout_total / Cres_total channels (Cin_total = the Cin the weights were packed with: the
 * tensor holds round_up(Cin,16) channel slots); the padding mode is whatever the halo of `in`
 * holds.  Odd k <= 15, Cout <= 32.  num_sms: CTAs of the persistent grid (<= 0: 148).
 *
 * Split precision (nsplit = 2): every BLK8 tensor is a pair of buffers x = hi + lo (lo = the fp16
 * rounding remainder of hi, written by the producers) and the weights are packed as W_hi and W_lo;
 * the kernel issues three MMAs per (chunk, tap, row): x_hi*W_hi + x_hi*W_lo + x_lo*W_hi, i.e. ~22
 * significand bits at 3x the tensor work.  nsplit = 1: single FP16 pass (11 bits, like TF32); the
 * *_lo pointers are NULL.
 * nsplit = 3: the two correction terms are evaluated by ONE e4m3 MMA of K = 32,
 * [e4m3(x) ; e4m3(x_lo*2^11)] * [e4m3(W_lo) ; e4m3(W*2^-11)], into the same accumulator: 2x the tensor
 * work of a single pass, ~15 significand bits.  The *_lo pointers are then the fp8 "q" buffers
 * (same byte geometry as a BLK8 buffer: 16 B per pixel per plane).
 * nsplit | PCNN_TC_SKIP_CORRECTION (mode 3 only): the tensors keep their e4m3 planes but THIS layer issues the
 * fp16 pass only (a layer whose rounding error does not matter pays single-pass cost inside a tc2 network).
 * out_halo_mode: PCNN_PAD_CONSTANT (0) leaves the halo of `out` alone; PCNN_PAD_SYMMETRIC makes the epilogue
 * also write the 7-wide mirrored ring (tf.pad SYMMETRIC for the next layer, fused; needs H, W >= 7). */
#define PCNN_TC_SKIP_CORRECTION 0x10
int pcnn_conv2d_tc(const void* in, const void* in_lo, const void* wpack, const float* bias,
                   const float* bn_scale, const float* bn_shift, const void* residual,
                   const void* residual_lo, const float* out_scale, void* out, void* out_lo, int B,
                   int Cin_total, int Cout, int Cout_total, int Cres_total, int H, int W, int k, int act,
                   int nsplit, float acc_scale, int out_halo_mode, int num_sms, void* stream);

/* Convolution of a SEPARABLE input on the row-group kernel ("row weights").  The DBCNN's first 2-D convolution
 * (models/Dirichlet_BC_NN_Legacy.py:155-160 of the reference) sees in[b,m,x,y] = h[b,m,y] * S[m,x] (the mode expansion of
 * :137-153; the two position channels are rank-1 too), so its row taps fold into per-row weights
 *     A_x[b_tap, m, co] = sum_a W[a, b_tap, m, co] * S[m, x + a - k/2]     (S zero outside [0,H): CONSTANT zero padding)
 * and out[b,co,x,y] = sum_{b_tap,m} A_x[b_tap,m,co] * h[b,m,y + b_tap - k/2]: ceil(Cin/16)*k MMAs per tile instead of
 * ceil(Cin/16)*k*(k+RT-1), and the [B,Cin,H,W] expansion is never written.
 * in_row: BLK8 fp16 tensor of the signals, [B][2*ceil(Cin/16)][1+14][W+14][8] (H = 1).  wrow: fp16
 * [ceil(Cin/16)][k][2][T][CP][8] with T = pcnn_conv_tc_rowweight_slots(Cout,k,H), CP = pcnn_conv_tc_channel_slots(Cout,k),
 * RT = 128/CP (5 for 24): slot t holds output row (t/RT)*RT + RT-1 - t%RT (zeros beyond H), K half = channels
 * 16c + 8*half + [0,8).  Single FP16 pass only; bias + activation in the epilogue; out as for pcnn_conv2d_tc. */
/* the signals of the DBCNN's separable layer as a BLK8 tensor with H = 1 (zero-initialised, pcnn_blk8_bytes(B, M+2, 1, n)):
 * channel m < M = h[b,m,y] * modew[b,m], channel M = 1, channel M+1 = posy[y] */
int pcnn_dbcnn_signal_blk8(const float* h, const float* modew, const float* posy, void* out, int B, int M, int n, void* stream);
int pcnn_conv_tc_rowweight_slots(int Cout, int k, int H);
int pcnn_conv2d_tc_rowweights(const void* in_row, const void* wrow, const float* bias, void* out, int B, int Cin,
                              int Cout, int Cout_total, int H, int W, int k, int act, float acc_scale, int num_sms,
                              void* stream);

/* ==== model-level API ==========================================================================================
 * The reference's unit of work is one Keras call: model([rhs, left, top, right, bottom, dx]) of Poisson_CNN_Legacy
 * (models/Poisson_CNN_Legacy.py:15-51), model([rhs, dx]) of Homogeneous_Poisson_NN_Legacy (:182-257) and
 * model([bc, dx, x_output_resolution]) of Dirichlet_BC_NN_Legacy_2 (models/Dirichlet_BC_NN_Legacy.py:124-166).  A handle
 * holds the parsed config, the name-addressed weights and their packed operand images; a forward call runs the whole layer
 * program (csrc/engine.cu) on the caller's stream inside ONE caller-provided workspace.
 *
 *   pcnn_create            config_json = {"hpnn_model": {...}, "dbcnn_model": {...}[, "jacobi_iterations": n]}: the sections
 *                          of the reference's experiment JSONs (experiments/pcnn_end_to_end.json; activations may stay the
 *                          "tf.nn.leaky_relu" strings).  "model" is accepted for "hpnn_model".  Either section alone gives a
 *                          single-network handle.  Config errors return PCNN_ERR_INVALID_ARGUMENT with the reference's
 *                          ValueError texts in pcnn_last_error().  One handle per (process, device); not thread-safe.
 *   pcnn_set_weight        one variable, Keras layout, from HOST memory (dtype 0 = float32, 1 = float64); names as in
 *                          poisson_cnn_b200/weights.py: "hpnn/pre_bottleneck/0/kernel", "dbcnn/boundary/3/resnet/conv1/bias",
 *                          "hpnn/final/2/resnet/bn0/gamma", ...  Allocates device memory (setup phase).
 *   pcnn_finalize_weights  checks that every variable of the config is present with the right shape, folds BatchNorm, packs
 *                          the tensor-core operand images.  precision: 0 strict FP32, 1 tc (single fp16 pass; NOT within the
 *                          2e-3 budget for the HPNN), 2 tc3, 3 tc2, 4 mixed (tc2 in the HPNN trunk, single pass in the DBCNN and
 *                          the HPNN's bottleneck branches: the default of the Python host).  May be called again.
 *   pcnn_*workspace_bytes  bytes a forward call of this shape needs (a dry run of the layer program with liveness-based
 *                          buffer reuse; batches larger than the micro-batch are processed in slices inside the same arena).
 *   pcnn_*forward          device pointers, fp32, dense: rhs/out [B,1,H,W]; left,right [B,1,W]; top,bottom [B,1,H]; dx [B,1];
 *                          bc [B,1,n] -> out [B,1,x_res,n].  The library allocates nothing and never synchronises here.  The
 *                          FIRST call for a (workspace pointer, shape) pair zero-fills the workspace and uploads the small
 *                          host-built tables (a host-blocking copy from pageable memory); later calls only launch kernels,
 *                          so a call is CUDA-graph capturable after one warm-up.  The workspace contents belong to the
 *                          handle between calls of the same shape (halo rings, tables).
 *   pcnn_set_microbatch    samples per slice (0 = automatic: 128 * 65536 / (H*W), the Python host's rule).
 *   pcnn_profile_conv_*    CUDA-event timing of every tensor-core conv launch with this (Cin, Cout, k) inside forward calls
 *                          (bench.py's live roofline measurement); _end synchronises on the recorded events. */
typedef struct pcnn_model* pcnn_handle;
enum { PCNN_PREC_FP32 = 0, PCNN_PREC_TC = 1, PCNN_PREC_TC3 = 2, PCNN_PREC_TC2 = 3, PCNN_PREC_MIXED = 4 };
int pcnn_create(const char* config_json, int device, pcnn_handle* out);
int pcnn_destroy(pcnn_handle handle);
int pcnn_set_weight(pcnn_handle handle, const char* name, const void* host_ptr, const int64_t* shape, int ndim, int dtype);
int pcnn_finalize_weights(pcnn_handle handle, int precision);
int pcnn_set_microbatch(pcnn_handle handle, int samples);
int pcnn_workspace_bytes(pcnn_handle handle, int B, int H, int W, size_t* bytes);
int pcnn_hpnn_workspace_bytes(pcnn_handle handle, int B, int H, int W, size_t* bytes);
int pcnn_dbcnn_workspace_bytes(pcnn_handle handle, int B, int n, int x_res, size_t* bytes);
int pcnn_hpnn_forward(pcnn_handle handle, const float* rhs, const float* dx, float* out, int B, int H, int W,
                      void* workspace, size_t workspace_bytes, void* stream);
int pcnn_dbcnn_forward(pcnn_handle handle, const float* bc, const float* dx, float* out, int B, int n, int x_res,
                       void* workspace, size_t workspace_bytes, void* stream);
int pcnn_forward(pcnn_handle handle, const float* rhs, const float* left, const float* top, const float* right,
                 const float* bottom, const float* dx, float* out, int B, int H, int W, void* workspace,
                 size_t workspace_bytes, void* stream);
/* Host-side table builders of the engine (no GPU involved), exported so that they can be pinned by CPU tests against the
 * Python host's tables: "pos" (a = n): cos(pi*linspace(0,1,n)); "sinh" (a = modes, b = x_res): the sinh basis of
 * models/Dirichlet_BC_NN_Legacy.py:106-112; "resize_idx" / "resize_w" (a = n_in, b = n_out, c = 0 nearest | 1 bilinear | 2
 * bicubic): the per-axis gather tables of tf.image.resize.  Return the number of elements written (or a negative status).
 * pcnn_host_rowweights: the fp16 operand image of pcnn_conv2d_tc_rowweights and its accumulator scale for a Keras kernel
 * [k,k,Cin,Cout] whose input channels are (Cin-2 sinh modes, posx, 1) on a grid of height x_res. */
long long pcnn_host_table(const char* what, int a, int b, int c, void* out, size_t out_bytes);
long long pcnn_host_rowweights(const float* kernel, int k, int Cin, int Cout, int x_res, void* out_img, size_t out_bytes,
                               float* acc_scale);
int pcnn_profile_conv_begin(pcnn_handle handle, int cin, int cout, int k, int max_launches);
int pcnn_profile_conv_end(pcnn_handle handle, int* launches, double* avg_ms, double* flops_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* PCNN_H_ */
