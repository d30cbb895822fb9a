/* A host in plain C for the model-level C ABI of libpcnn.so (include/pcnn.h): what a non-Python caller of
 * Poisson_CNN_Legacy.call (reference: poisson_CNN/models/Poisson_CNN_Legacy.py:15-51) looks like.
 *
 *   pcnn_host <config.json> <weights.bin> <inputs.bin> <B> <H> <W> <precision 0..4> <out.bin>
 *
 * config.json  the reference's experiment JSON (sections hpnn_model / dbcnn_model; other keys are ignored)
 * weights.bin  records { uint32 name_len; char name[name_len]; uint32 ndim; int64 shape[ndim]; float data[prod(shape)] }
 * inputs.bin   float32: rhs[B*H*W], left[B*W], top[B*H], right[B*W], bottom[B*H], dx[B]
 * out.bin      float32: prediction[B*H*W]
 * Only libpcnn.so and the CUDA runtime are linked: no Python, no torch. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pcnn.h"

#define CK(call)                                                                          \
    do {                                                                                  \
        int st_ = (call);                                                                 \
        if (st_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, st_, pcnn_last_error()); return 1; } \
    } while (0)
#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } \
    } while (0)

static char* read_all(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); return NULL; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)sz + 1);
    if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return NULL; }
    buf[sz] = 0;
    fclose(f);
    if (n) *n = (size_t)sz;
    return buf;
}

int main(int argc, char** argv) {
    if (argc != 9) { fprintf(stderr, "usage: %s config.json weights.bin inputs.bin B H W precision out.bin\n", argv[0]); return 2; }
    const int B = atoi(argv[4]), H = atoi(argv[5]), W = atoi(argv[6]), precision = atoi(argv[7]);
    size_t n = 0;
    char* cfg = read_all(argv[1], &n);
    if (!cfg) return 1;
    pcnn_handle h = NULL;
    CK(pcnn_create(cfg, 0, &h));
    free(cfg);

    char* wb = read_all(argv[2], &n);
    if (!wb) return 1;
    size_t off = 0;
    int nvars = 0;
    while (off < n) {
        uint32_t len, ndim;
        memcpy(&len, wb + off, 4); off += 4;
        char name[512];
        if (len >= sizeof(name)) { fprintf(stderr, "variable name too long\n"); return 1; }
        memcpy(name, wb + off, len); name[len] = 0; off += len;
        memcpy(&ndim, wb + off, 4); off += 4;
        int64_t shape[8];
        size_t numel = 1;
        for (uint32_t i = 0; i < ndim; ++i) { memcpy(&shape[i], wb + off, 8); off += 8; numel *= (size_t)shape[i]; }
        CK(pcnn_set_weight(h, name, wb + off, shape, (int)ndim, 0));
        off += numel * 4;
        ++nvars;
    }
    free(wb);
    CK(pcnn_finalize_weights(h, precision));

    const size_t plane = (size_t)H * W, n_in = (size_t)B * (plane + 2 * W + 2 * H + 1);
    size_t got = 0;
    float* host = (float*)read_all(argv[3], &got);
    if (!host || got != n_in * 4) { fprintf(stderr, "inputs.bin: expected %zu bytes, got %zu\n", n_in * 4, got); return 1; }
    float *d_in = NULL, *d_out = NULL;
    void* ws = NULL;
    size_t ws_bytes = 0;
    CK(pcnn_workspace_bytes(h, B, H, W, &ws_bytes));
    CU(cudaMalloc((void**)&d_in, n_in * 4));
    CU(cudaMalloc((void**)&d_out, (size_t)B * plane * 4));
    CU(cudaMalloc(&ws, ws_bytes));
    CU(cudaMemcpy(d_in, host, n_in * 4, cudaMemcpyHostToDevice));
    const float* rhs = d_in;
    const float* left = rhs + (size_t)B * plane;
    const float* top = left + (size_t)B * W;
    const float* right = top + (size_t)B * H;
    const float* bottom = right + (size_t)B * W;
    const float* dx = bottom + (size_t)B * H;
    cudaStream_t st;
    CU(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CK(pcnn_forward(h, rhs, left, top, right, bottom, dx, d_out, B, H, W, ws, ws_bytes, st));      /* prepares the workspace */
    CU(cudaEventRecord(e0, st));
    CK(pcnn_forward(h, rhs, left, top, right, bottom, dx, d_out, B, H, W, ws, ws_bytes, st));
    CU(cudaEventRecord(e1, st));
    CU(cudaStreamSynchronize(st));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    float* out = (float*)malloc((size_t)B * plane * 4);
    CU(cudaMemcpy(out, d_out, (size_t)B * plane * 4, cudaMemcpyDeviceToHost));
    FILE* f = fopen(argv[8], "wb");
    if (!f || fwrite(out, 4, (size_t)B * plane, f) != (size_t)B * plane) { perror(argv[8]); return 1; }
    fclose(f);
    printf("pcnn_host: %d variables, workspace %.1f MB, forward of %d x %dx%d (precision %d): %.3f ms, %lld kernel launches so far\n",
           nvars, ws_bytes / 1e6, B, H, W, precision, ms, pcnn_launch_count());
    CK(pcnn_destroy(h));
    cudaFree(ws); cudaFree(d_in); cudaFree(d_out);
    free(out); free(host);
    return 0;
}
