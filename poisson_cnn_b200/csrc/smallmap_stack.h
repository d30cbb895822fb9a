// Internal interface of csrc/smallmap_stack.cu for the engine: several small-map layer programs (one per bottleneck branch)
// in ONE launch.  The C-ABI export pcnn_smallmap_stack_f32 is the same call with a single program.
#pragma once
#include <cuda_runtime.h>

namespace pcnn {
namespace sms {

constexpr int MAX_PROGRAMS = 8;      // programs per launch (the shipped HPNN has 3 small-map branches at 256^2, 5 at 64^2)

// one fused conv + resnet chain on maps of H*W <= 64 pixels: the arguments of pcnn_smallmap_stack_f32 (include/pcnn.h)
struct StackDesc {
    const float* in;
    float* out;
    int H, W, Cin0, n_layers;
    const float* const* kernels;
    const float* const* biases;
    const float* const* bn_scale;
    const float* const* bn_shift;
    const int* ksize;
    const int* cin;
    const int* cout;
    const int* flags;
    int act, pad_mode;
    float pad_value;
};

// runs n (<= MAX_PROGRAMS) programs over the same batch B as grid (B, n): the CTAs of different programs run side by side
// instead of in n dependent launches of B CTAs each (the programs are latency-bound: ~0.35 ms for one sample)
int smallmap_stack_multi(const StackDesc* descs, int n, int B, cudaStream_t stream);

}  // namespace sms
}  // namespace pcnn
