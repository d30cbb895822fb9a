// Handle-level C ABI of the hot path: pcnn_create / pcnn_set_weight / pcnn_finalize_weights / pcnn_workspace_bytes /
// pcnn_hpnn_forward / pcnn_dbcnn_forward / pcnn_forward  (SURVEY.md 8(b) "C exports").
//
// The reference's unit of work is model([rhs, left, top, right, bottom, dx]) (poisson_CNN/models/Poisson_CNN_Legacy.py:15-51,
// Homogeneous_Poisson_NN_Legacy.py:182-257, Dirichlet_BC_NN_Legacy.py:124-166).  This file holds the LAYER PROGRAM of those
// three calls in C++: config parsing, name-addressed weights, BatchNorm folding, operand packing for the tensor-core
// kernels, the host-built tables (tf.image.resize gather tables, cos/sinh bases, SPP boxes, separable row weights), a
// liveness-based activation arena inside ONE caller-provided workspace, and the ~165 kernel launches of a forward pass
// (all through the per-op entry points of this library, on the caller's stream).  No allocation and no synchronisation
// happens inside a forward call; the first call for a (workspace, shape) pair zero-fills the workspace and uploads the
// tables (a host-blocking copy, once), every later call only launches kernels -- CUDA-graph capturable after one warm-up.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_fp16.h>

#include "engine_json.h"
#include "smallmap_stack.h"
#include "pcnn_common.cuh"

namespace pcnn {
namespace eng {

#define TRY(expr)                        \
    do {                                 \
        const int _st = (expr);          \
        if (_st != PCNN_OK) return _st;  \
    } while (0)
// kernel launches are skipped in a dry run (workspace sizing replays the program without touching the device)
#define RUN(c, expr)                     \
    do {                                 \
        if (!(c).dry) TRY(expr);         \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

enum { PREC_FP32 = 0, PREC_TC = 1, PREC_TC3 = 2, PREC_TC2 = 3, PREC_MIXED = 4 };
enum { RESIZE_NEAREST = 0, RESIZE_BILINEAR = 1, RESIZE_BICUBIC = 2 };

// ------------------------------------------------------------------------------------------------ config
static int act_enum(const std::string& a) {
    if (a.empty() || a == "None" || a == "linear" || a.find("linear") != std::string::npos) return PCNN_ACT_LINEAR;
    if (a.find("leaky_relu") != std::string::npos) return PCNN_ACT_LEAKY_RELU;
    if (a.find("tanh") != std::string::npos) return PCNN_ACT_TANH;
    throw std::runtime_error("unsupported activation: " + a + " (supported: leaky_relu, tanh, linear)");
}
static int pad_enum(std::string m) {
    for (auto& ch : m) ch = (char)std::toupper((unsigned char)ch);
    if (m == "CONSTANT") return PCNN_PAD_CONSTANT;
    if (m == "SYMMETRIC") return PCNN_PAD_SYMMETRIC;
    if (m == "REFLECT") return PCNN_PAD_REFLECT;
    throw std::runtime_error("unsupported padding mode " + m);
}
static int resize_enum(std::string m) {
    for (auto& ch : m) ch = (char)std::tolower((unsigned char)ch);
    if (m == "nearest") return RESIZE_NEAREST;
    if (m == "bilinear") return RESIZE_BILINEAR;
    if (m == "bicubic") return RESIZE_BICUBIC;
    throw std::runtime_error("unsupported resize method " + m + " (nearest, bilinear, bicubic)");
}

struct BlockCfg {     // one resolution branch (blocks/bottleneck_block.py:8-118)
    bool deconv = true;
    int index = 0, ds = 1, us = 1, ksize = 1, n_convs = 1, act = 0, deconv_act = 0, pad = 0, resize_method = RESIZE_BILINEAR;
    float pad_value = 0.f;
    bool use_bn = false;
    std::string name;
};

struct StackCfg {     // a list of conv stages: filters, kernel sizes, padding, activation
    std::vector<int> filters, ksizes;
    int pad = 0, act = 0, nreg = 2;
    float pad_value = 0.f;
};

struct HpnnCfg {
    bool present = false, use_bn = false, use_pos = true, use_scaling = false;
    int bc_type = PCNN_BC_DIRICHLET, postsmooth = 0, F = 0;
    StackCfg pre, fin;
    std::vector<BlockCfg> blocks;     // deconv branches (descending downsampling factor), then multilinear branches
    int sc_stages = 2, sc_ratio = 2, sc_act = 0, sc_ksize = 3, sc_filters = 0;
    std::vector<std::vector<int>> sc_levels;
};

struct DbcnnCfg {
    bool present = false, use_bn = false;
    StackCfg bnd, fin;
    std::vector<std::vector<int>> spp_levels;
    int spp_mode = PCNN_POOL_AVG, postsmooth = 0, nmodes = 0;
    std::vector<int> mlp_units, mlp_acts;
};

static std::vector<std::vector<int>> parse_levels(const json::Value& v) {
    std::vector<std::vector<int>> out;
    for (const auto& e : v.arr) {
        std::vector<int> lv;
        if (e.type == json::Value::Num) lv.push_back((int)e.num);
        else for (const auto& q : e.arr) lv.push_back((int)q.num);
        out.push_back(lv);
    }
    return out;
}

static StackCfg parse_stack(const json::Value& v, const char* what) {
    StackCfg s;
    s.filters = v.ints("filters");
    s.ksizes = v.ints("kernel_sizes");
    if (s.filters.empty() || s.ksizes.size() < s.filters.size()) throw std::runtime_error(std::string(what) + ": filters / kernel_sizes mismatch");
    s.pad = pad_enum(v.string("padding_mode", "CONSTANT"));
    s.pad_value = (float)v.number("constant_padding_value", 0.0);
    s.act = act_enum(v.string("activation", ""));
    s.nreg = (int)v.number("final_regular_conv_stages", 2);
    return s;
}

static HpnnCfg parse_hpnn(const json::Value& v) {
    HpnnCfg h;
    h.present = true;
    if (v.string("data_format", "channels_first") != "channels_first") throw std::runtime_error("the CUDA path is channels_first (every shipped config)");
    h.use_bn = v.boolean("use_batchnorm", false);
    h.use_pos = v.boolean("use_positional_embeddings", true);
    h.use_scaling = v.boolean("use_scaling", false);
    h.postsmooth = (int)v.number("postsmoother_iterations", 5);
    std::string bc = v.string("bc_type", "dirichlet");
    for (auto& ch : bc) ch = (char)std::tolower((unsigned char)ch);
    if (bc != "dirichlet" && bc != "neumann") throw std::runtime_error("bc_type can only be neumann or dirichlet.");
    h.bc_type = bc == "dirichlet" ? PCNN_BC_DIRICHLET : PCNN_BC_NEUMANN;
    if (!v.has("pre_bottleneck_convolutions_config")) throw std::runtime_error("Provide a config for pre bottleneck convolutions");
    if (!v.has("bottleneck_deconv_config") || !v.has("bottleneck_multilinear_config")) throw std::runtime_error("Provide a config for bottleneck blocks");
    if (!v.has("final_convolutions_config")) throw std::runtime_error("Provide a config for final convolutions");
    if (h.use_scaling && !v.has("scaling_config")) throw std::runtime_error("use_scaling=True needs a scaling_config");
    h.pre = parse_stack(v.at("pre_bottleneck_convolutions_config"), "pre_bottleneck_convolutions_config");
    h.fin = parse_stack(v.at("final_convolutions_config"), "final_convolutions_config");
    const json::Value& dc = v.at("bottleneck_deconv_config");
    const json::Value& mc = v.at("bottleneck_multilinear_config");
    h.F = (int)dc.number("filters", 0);
    if (h.F <= 0 || (int)mc.number("filters", 0) != h.F) throw std::runtime_error("bottleneck configs must share one positive filter count");
    for (int kind = 0; kind < 2; ++kind) {
        const json::Value& c = kind == 0 ? dc : mc;
        std::string m = c.string("downsampling_method", "conv");
        for (auto& ch : m) ch = (char)std::tolower((unsigned char)ch);
        if (m != "conv" && m != "pool") throw std::runtime_error("Downsampling method can only be conv or pool");
        std::string pm = c.string("pool_downsampling_method", "max");
        for (auto& ch : pm) ch = (char)std::tolower((unsigned char)ch);
        if (m != "pool" || !c.boolean("use_resnet", false)) throw std::runtime_error("only downsampling_method='pool' with use_resnet=True (every shipped config) is built");
        if (pm != "average" && pm != "avg") throw std::runtime_error("only average pooling (every shipped config) is built");
        const std::vector<int> ds = c.ints("downsampling_factors");
        const std::vector<int> us = c.has("upsampling_factors") ? c.ints("upsampling_factors") : ds;
        const std::vector<int> ks = c.ints("conv_kernel_sizes"), nc = c.ints("n_convs");
        std::vector<BlockCfg> blocks;
        for (size_t i = 0; i < ds.size(); ++i) {
            BlockCfg b;
            b.deconv = kind == 0;
            b.index = (int)i;
            b.ds = ds[i];
            b.us = us[i];
            b.ksize = ks.at(i);
            b.n_convs = nc.at(i);
            b.act = act_enum(c.string("conv_activation", ""));
            b.deconv_act = act_enum(c.string("deconv_activation", ""));
            b.pad = pad_enum(c.string("padding_mode", "constant"));
            b.pad_value = (float)c.number("constant_padding_value", 0.0);
            b.use_bn = h.use_bn;
            if (kind == 0) {
                if (c.ints("deconv_kernel_sizes").at(i) != b.us) throw std::runtime_error("deconvupscale: only kernel size == stride (every shipped config) is built");
            } else if (c.has("resize_methods")) {
                b.resize_method = resize_enum(c.at("resize_methods").arr.at(i).str);
            } else {
                b.resize_method = resize_enum(c.string("resize_method", "bilinear"));
            }
            b.name = std::string(kind == 0 ? "bottleneck_deconv/" : "bottleneck_multilinear/") + std::to_string(i);
            blocks.push_back(b);
        }
        std::stable_sort(blocks.begin(), blocks.end(), [](const BlockCfg& a, const BlockCfg& b) { return a.ds > b.ds; });
        h.blocks.insert(h.blocks.end(), blocks.begin(), blocks.end());
    }
    if (h.use_scaling) {
        const json::Value& sc = v.at("scaling_config");
        h.sc_stages = (int)sc.number("stages", 2);
        h.sc_ratio = (int)sc.number("downsampling_ratio_per_stage", 2);
        h.sc_act = act_enum(sc.string("activation", ""));
        h.sc_ksize = (int)sc.number("kernel_size", 3);
        h.sc_filters = (int)sc.number("filters", 0);
        if (h.sc_ksize % 2 == 0) throw std::runtime_error("Scaling: even kernel sizes with Keras 'same' padding are not built");
        if (sc.has("spp_levels")) h.sc_levels = parse_levels(sc.at("spp_levels"));
        else h.sc_levels = {{2, 2}, {3}, {5}};
    }
    return h;
}

static DbcnnCfg parse_dbcnn(const json::Value& v) {
    DbcnnCfg d;
    d.present = true;
    if (v.string("data_format", "channels_first") != "channels_first") throw std::runtime_error("the CUDA path is channels_first (every shipped config)");
    d.use_bn = v.boolean("use_batchnorm", false);
    d.postsmooth = (int)v.number("postsmoother_iterations", 0);
    if (!v.has("boundary_conv_config")) throw std::runtime_error("Provide a config for the boundary convolutions.");
    if (!v.has("spp_config")) throw std::runtime_error("Provide a config for the Spatial Pyramid Pooling.");
    if (!v.has("final_convolutions_config")) throw std::runtime_error("Provide a config for the domain convolutions.");
    if (!v.has("domain_info_mlp_config")) throw std::runtime_error("Provide a config for the domain info MLP.");
    d.bnd = parse_stack(v.at("boundary_conv_config"), "boundary_conv_config");
    d.fin = parse_stack(v.at("final_convolutions_config"), "final_convolutions_config");
    const json::Value& spp = v.at("spp_config");
    d.spp_levels = parse_levels(spp.at("levels"));
    std::string pt = spp.string("pooling_type", "average");
    for (auto& ch : pt) ch = (char)std::tolower((unsigned char)ch);
    if (pt == "average" || pt == "avg") d.spp_mode = PCNN_POOL_AVG;
    else if (pt == "max") d.spp_mode = PCNN_POOL_MAX;
    else throw std::runtime_error("unknown SPP pooling_type " + pt);
    const json::Value& mlp = v.at("domain_info_mlp_config");
    d.mlp_units = mlp.ints("units");
    for (const auto& a : mlp.at("activations").arr) d.mlp_acts.push_back(act_enum(a.type == json::Value::Str ? a.str : ""));
    if (d.mlp_units.empty() || d.mlp_acts.size() != d.mlp_units.size()) throw std::runtime_error("domain_info_mlp_config: units / activations mismatch");
    d.nmodes = d.mlp_units.back();
    if (d.bnd.filters.back() != d.nmodes) throw std::runtime_error("boundary_conv_config filters[-1] must equal domain_info_mlp_config units[-1]");
    return d;
}

static int spp_bins(const std::vector<std::vector<int>>& levels, int ndims) {
    int n = 0;
    for (const auto& lv : levels) {
        if (lv.size() == 1) { int p = 1; for (int i = 0; i < ndims; ++i) p *= lv[0]; n += p; }
        else { int p = 1; for (int v : lv) p *= v; n += p; }
    }
    return n;
}

// ------------------------------------------------------------------------------------------------ host-built tables
// tf.linspace(0., 1., n) in float32: start + i*step, last element := stop
static std::vector<float> linspace01(int n) {
    std::vector<float> v((size_t)n, 0.f);
    if (n == 1) return v;
    const float step = 1.0f / (float)(n - 1);
    for (int i = 0; i < n; ++i) v[i] = (float)i * step;
    v[n - 1] = 1.0f;
    return v;
}
static const float kPiF = 3.14159274101257324f;      // np.float32(math.pi)

// cos(pi * linspace(0,1,n)) in float32 (generate_position_embeddings, Homogeneous_Poisson_NN_Legacy.py:172-180)
static std::vector<float> position_table(int n) {
    std::vector<float> v = linspace01(n);
    for (auto& x : v) x = (float)std::cos((double)(kPiF * x));
    return v;
}

// build_series_x_dir_components (Dirichlet_BC_NN_Legacy.py:106-112): sinh(m pi (x - 1)) / max|.| per mode, float32
static std::vector<float> sinh_basis(int M, int xres) {
    const std::vector<float> xbar = linspace01(xres);
    std::vector<float> s((size_t)M * xres);
    for (int m = 0; m < M; ++m) {
        float mx = 0.f;
        for (int x = 0; x < xres; ++x) {
            const float arg = (float)(m + 1) * (kPiF * (xbar[x] - 1.0f));
            const float v = (float)std::sinh((double)arg);
            s[(size_t)m * xres + x] = v;
            mx = std::max(mx, std::fabs(v));
        }
        const float inv = 1.0f / mx;
        for (int x = 0; x < xres; ++x) s[(size_t)m * xres + x] *= inv;
    }
    return s;
}

struct AxisTable { int taps = 0; std::vector<int32_t> idx; std::vector<float> w; };

static const std::vector<float>& bicubic_table() {     // TF resize_bicubic: 1024-step Keys (a = -0.5) table, float32
    static std::vector<float> tab;
    if (tab.empty()) {
        const int n = 1024;
        const float a = -0.5f;
        tab.resize(2 * (n + 1));
        for (int i = 0; i <= n; ++i) {
            const float x = (float)i / (float)n;
            tab[2 * i] = ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
            const float x1 = x + 1.0f;
            tab[2 * i + 1] = ((a * x1 - 5.0f * a) * x1 + 8.0f * a) * x1 - 4.0f * a;
        }
    }
    return tab;
}

// per-axis gather indices / weights of tf.image.resize (half-pixel centres, antialias=False); layers/Upsample.py:56-59
static AxisTable resize_axis_table(int n_in, int n_out, int method) {
    AxisTable t;
    const float scale = (float)n_in / (float)n_out;
    if (method == RESIZE_NEAREST) {
        t.taps = 1;
        for (int o = 0; o < n_out; ++o) {
            long long i = (long long)std::floor(((float)o + 0.5f) * scale);
            t.idx.push_back((int32_t)std::min<long long>(std::max<long long>(i, 0), n_in - 1));
            t.w.push_back(1.0f);
        }
        return t;
    }
    for (int o = 0; o < n_out; ++o) {
        const float src = ((float)o + 0.5f) * scale - 0.5f;
        const float fl = std::floor(src);
        if (method == RESIZE_BILINEAR) {
            t.taps = 2;
            const long long lo = std::max<long long>((long long)fl, 0), hi = std::min<long long>((long long)std::ceil(src), n_in - 1);
            const float lerp = src - fl;
            t.idx.push_back((int32_t)lo); t.idx.push_back((int32_t)hi);
            t.w.push_back(1.0f - lerp); t.w.push_back(lerp);
        } else {
            t.taps = 4;
            const std::vector<float>& tab = bicubic_table();
            const int n = 1024;
            const long long loc = (long long)fl;
            const long long off = (long long)std::nearbyint((src - fl) * (float)n);
            float w[4] = {tab[off * 2 + 1], tab[off * 2], tab[(n - off) * 2], tab[(n - off) * 2 + 1]};
            long long raw[4] = {loc - 1, loc, loc + 1, loc + 2};
            float sum = 0.f;
            int32_t idx[4];
            for (int a = 0; a < 4; ++a) {
                const long long cl = std::min<long long>(std::max<long long>(raw[a], 0), n_in - 1);
                idx[a] = (int32_t)cl;
                if (cl != raw[a]) w[a] = 0.f;
            }
            for (int a = 0; a < 4; ++a) sum += w[a];
            if (std::fabs(sum) >= 1000.0f * 1.17549435e-38f) {
                const float inv = 1.0f / sum;
                for (int a = 0; a < 4; ++a) w[a] *= inv;
            }
            for (int a = 0; a < 4; ++a) { t.idx.push_back(idx[a]); t.w.push_back(w[a]); }
        }
    }
    return t;
}

// dataset/utils/split_indices.py:4-26 (numpy.array_split boundaries)
static std::vector<int> split_indices(int n, int sections) {
    const int per = n / sections, extra = n % sections;
    std::vector<int> e{0};
    for (int i = 0; i < sections; ++i) e.push_back(e.back() + per + (i < extra ? 1 : 0));
    return e;
}

// bin boxes (y0,y1,x0,x1) in the order SpatialPyramidPool.call emits them (layers/SpatialPyramidPool.py:48-66)
static std::vector<int32_t> spp_boxes(const std::vector<std::vector<int>>& levels, int H, int W, int ndims) {
    std::vector<int32_t> b;
    for (const auto& lv0 : levels) {
        std::vector<int> lv = lv0;
        if (lv.size() == 1) lv = std::vector<int>((size_t)ndims, lv0[0]);
        else if ((int)lv.size() != ndims) throw std::runtime_error("Each SPP level must have a pool size with ndims or 1 element(s).");
        if (ndims == 1) {
            const std::vector<int> ex = split_indices(W, lv[0]);
            for (int i = 0; i < lv[0]; ++i) { b.push_back(0); b.push_back(H); b.push_back(ex[i]); b.push_back(ex[i + 1]); }
        } else {
            const std::vector<int> ey = split_indices(H, lv[0]), ex = split_indices(W, lv[1]);
            for (int i = 0; i < lv[0]; ++i)
                for (int j = 0; j < lv[1]; ++j) { b.push_back(ey[i]); b.push_back(ey[i + 1]); b.push_back(ex[j]); b.push_back(ex[j + 1]); }
        }
    }
    return b;
}

static float pow2_prescale(float amax) {     // power of two that puts max|W| near 2^9 (fp16 hi and lo parts both normal)
    if (!(amax > 0.f) || !std::isfinite(amax)) return 1.0f;
    const int e = std::max(-24, std::min(24, (int)std::floor(std::log2(512.0 / (double)amax))));
    return (float)std::ldexp(1.0, e);
}

// ------------------------------------------------------------------------------------------------ weights
struct Weight {
    std::vector<int64_t> shape;
    std::vector<float> host;
    float* dev = nullptr;
    size_t numel() const { return host.size(); }
};

struct ConvW {      // fp32 convolution operands (Keras layouts) + optional folded BatchNorm
    const float* kernel = nullptr;
    const float* bias = nullptr;
    const float* bn_scale = nullptr;
    const float* bn_shift = nullptr;
    int k = 0, cin = 0, cout = 0;
    const Weight* w = nullptr;
};

struct TcPack { void* packed = nullptr; int k = 0, cin = 0, cout = 0, nsplit = 1; float acc_scale = 1.f; };

struct StackLayers {       // layer program of the fused 1-D / small-map stack kernels
    std::vector<const float*> kernels, biases, bn_scale, bn_shift;
    std::vector<int> ksize, cin, cout, flags;
    int n() const { return (int)ksize.size(); }
};

// ------------------------------------------------------------------------------------------------ arena
struct Halo { int mode = PCNN_PAD_CONSTANT; int pad = 7; bool operator!=(const Halo& o) const { return mode != o.mode || pad != o.pad; } };

struct SlotKey {
    int kind, a, b, c, d, e, f;      // kind: 0 raw bytes, 1 BLK8 hi, 2 BLK8 lo; the rest: shape / role / flags
    bool operator==(const SlotKey& o) const { return kind == o.kind && a == o.a && b == o.b && c == o.c && d == o.d && e == o.e && f == o.f; }
};

struct Slot { SlotKey key; size_t off = 0, bytes = 0; bool busy = false; Halo halo; };

struct TableRec { int slot; };

struct Prepared {       // what a workspace currently holds: one shape's slots, their halo states and the uploaded tables
    std::string shape_key;
    std::vector<Slot> slots;
    size_t top = 0;
    std::map<std::string, int> tables;      // table name -> slot
    std::map<std::string, float> scalars;   // host-side companions of tables (e.g. the pre-scale of the row weights)
    bool valid = false;
    void reset() { for (auto& s : slots) if (s.key.kind != 3) s.busy = false; }      // tables stay, activations restart
};

struct F32 {       // NCHW fp32 tensor (possibly a channel slice of a wider buffer: bs = batch stride in elements)
    int slot = -1;
    float* p = nullptr;
    int B = 0, C = 0, H = 0, W = 0;
    long long bs = 0;
};

struct B8 {        // BLK8 fp16 tensor (+ second buffer in split precision)
    int s_hi = -1, s_lo = -1;
    char* hi = nullptr;
    char* lo = nullptr;
    int mode = 1, B = 0, C = 0, H = 0, W = 0;
    Halo halo;
    bool live() const { return s_hi >= 0; }
    char* plane(char* base, int c_offset) const { return base ? base + (size_t)(c_offset / 8) * (H + 14) * (W + 14) * 16 : nullptr; }
};

struct Model;

struct Ctx {
    Model* m = nullptr;
    bool dry = false;
    char* base = nullptr;
    cudaStream_t st = nullptr;
    Prepared* prep = nullptr;
    int num_sms = 148;

    int alloc(const SlotKey& key, size_t bytes) {
        bytes = align_up(std::max<size_t>(bytes, 16), 1024);
        auto& slots = prep->slots;
        for (size_t i = 0; i < slots.size(); ++i)
            if (!slots[i].busy && slots[i].key == key && slots[i].bytes == bytes) { slots[i].busy = true; return (int)i; }
        Slot s;
        s.key = key; s.off = prep->top; s.bytes = bytes; s.busy = true;
        prep->top += bytes;
        slots.push_back(s);
        return (int)slots.size() - 1;
    }
    void release(int slot) { if (slot >= 0) prep->slots[slot].busy = false; }
    char* ptr(int slot) const { return dry ? nullptr : base + prep->slots[slot].off; }

    F32 f32(int Bcap, int B, int C, int H, int W) {
        F32 t;
        t.slot = alloc(SlotKey{0, 0, 0, 0, 0, 0, 0}, (size_t)Bcap * C * H * W * sizeof(float));
        t.p = reinterpret_cast<float*>(ptr(t.slot));
        t.B = B; t.C = C; t.H = H; t.W = W; t.bs = (long long)C * H * W;
        return t;
    }
    float* vec(size_t n) { return reinterpret_cast<float*>(ptr(alloc(SlotKey{0, 0, 0, 0, 0, 0, 0}, n * sizeof(float)))); }
    int vec_slot(size_t bytes) { return alloc(SlotKey{0, 0, 0, 0, 0, 0, 0}, bytes); }
    void free(F32& t) { release(t.slot); t.slot = -1; }

    B8 b8(int Bcap, int B, int C, int H, int W, int mode, bool sym) {
        B8 t;
        // Slots are shared by BYTE GEOMETRY (Cpad = round_up(C, 16) channel slots), not by the exact channel count: every
        // producer rewrites all ceil(C/8) live planes as whole 16-byte pixels (padding channels of the last live plane come
        // out as exact zeros) and no consumer reads a plane beyond ceil(Cin/8) (tap pairing reads the lone last plane
        // only), so stale data of a previous tenant is either overwritten or never touched.  The halo ring is what must
        // match, hence the `sym` bit in the key.  32/28/24/20-channel tensors of the trunk share three slots instead of
        // twelve (59.6 GB -> see tests/test_gpu_engine.py for the planned workspace at the headline shape).
        const size_t bytes = pcnn_blk8_bytes(Bcap, C, H, W);
        static const bool exact = std::getenv("PCNN_ENGINE_EXACT_SLOTS") != nullptr;      // debugging aid: one class per C
        const int cpad8 = exact ? 1000 + C : (C + 15) / 16 * 2;
        t.s_hi = alloc(SlotKey{1, Bcap, cpad8, H, W, 0, sym ? 1 : 0}, bytes);
        t.hi = ptr(t.s_hi);
        t.halo = prep->slots[t.s_hi].halo;
        if (mode >= 2) {
            t.s_lo = alloc(SlotKey{2, Bcap, cpad8, H, W, mode, sym ? 1 : 0}, bytes);
            t.lo = ptr(t.s_lo);
            const Halo hl = prep->slots[t.s_lo].halo;
            if (hl != t.halo) { t.halo.mode = t.halo.mode != PCNN_PAD_CONSTANT ? t.halo.mode : hl.mode; t.halo.pad = -1; }
        }
        t.mode = mode; t.B = B; t.C = C; t.H = H; t.W = W;
        return t;
    }
    void free(B8& t) {
        if (t.s_hi >= 0) { prep->slots[t.s_hi].halo = t.halo; release(t.s_hi); }
        if (t.s_lo >= 0) { prep->slots[t.s_lo].halo = t.halo; release(t.s_lo); }
        t.s_hi = t.s_lo = -1;
    }
};

static bool fusable_halo(int mode, int H, int W) { return mode == PCNN_PAD_SYMMETRIC && H >= 7 && W >= 7; }

// ------------------------------------------------------------------------------------------------ the model
struct Model {
    int device = 0, num_sms = 148;
    HpnnCfg hp;
    DbcnnCfg db;
    std::unordered_map<std::string, Weight> w;
    std::unordered_map<std::string, std::pair<float*, float*>> bn;     // folded (scale, shift)
    std::map<std::pair<std::string, int>, TcPack> tc;                  // (layer, nsplit) -> packed operand image
    std::map<std::string, void*> deconv_tc, deconv_f32;                // packed deconv kernels for the fused upsample-merge
    std::vector<void*> owned;                                          // every cudaMalloc of this handle
    int precision = -1, hp_mode = 0, db_mode = 0;
    bool branch_single = false, finalized = false;
    int microbatch = 0;                                                // 0: 128 * 65536 / (H*W) like the Python host
    std::map<void*, Prepared> prepared;
    std::map<std::string, size_t> need_cache;
    // optional CUDA-event timing of one conv shape (bench.py's live roofline measurement)
    int prof_cin = 0, prof_cout = 0, prof_k = 0;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    double prof_flops = 0.0;

    ~Model() {
        for (void* p : owned) cudaFree(p);
        for (auto e : prof_ev) cudaEventDestroy(e);
    }

    int dev_alloc(void** p, size_t bytes) {
        PCNN_CHECK_CUDA(cudaMalloc(p, std::max<size_t>(bytes, 16)));
        owned.push_back(*p);
        return PCNN_OK;
    }

    const Weight* find(const std::string& name) const {
        auto it = w.find(name);
        return it == w.end() ? nullptr : &it->second;
    }

    int conv(const std::string& name, int nd, ConvW* out) const {
        const Weight* k = find(name + "/kernel");
        if (!k) { set_error("weight '%s/kernel' was never set", name.c_str()); return PCNN_ERR_INVALID_ARGUMENT; }
        if ((int)k->shape.size() != nd + 2) { set_error("weight '%s/kernel' has rank %zu, expected %d", name.c_str(), k->shape.size(), nd + 2); return PCNN_ERR_INVALID_ARGUMENT; }
        out->w = k;
        out->kernel = k->dev;
        out->k = (int)k->shape[0];
        out->cin = (int)k->shape[nd];
        out->cout = (int)k->shape[nd + 1];
        const Weight* b = find(name + "/bias");
        out->bias = b ? b->dev : nullptr;
        out->bn_scale = out->bn_shift = nullptr;
        return PCNN_OK;
    }
    void with_bn(const std::string& name, ConvW* c) const {
        auto it = bn.find(name);
        if (it != bn.end()) { c->bn_scale = it->second.first; c->bn_shift = it->second.second; }
    }
};

// shape checks of the weights a config needs happen in finalize (every layer descriptor is resolved once there)
static int check_shape(const Model& m, const std::string& name, std::initializer_list<int64_t> shape) {
    const Weight* w = m.find(name);
    if (!w) { set_error("finalize_weights: variable '%s' missing", name.c_str()); return PCNN_ERR_INVALID_ARGUMENT; }
    if (w->shape != std::vector<int64_t>(shape)) {
        std::string got, want;
        for (auto v : w->shape) got += std::to_string(v) + ",";
        for (auto v : shape) want += std::to_string(v) + ",";
        set_error("finalize_weights: '%s' has shape [%s], expected [%s]", name.c_str(), got.c_str(), want.c_str());
        return PCNN_ERR_INVALID_ARGUMENT;
    }
    return PCNN_OK;
}

struct Expect { Model& m; std::string prefix; bool missing_bias_ok; int st = PCNN_OK;
    void conv(const std::string& n, int nd, int k, int cin, int cout) {
        if (st) return;
        if (nd == 2) st = check_shape(m, prefix + n + "/kernel", {k, k, cin, cout});
        else st = check_shape(m, prefix + n + "/kernel", {k, cin, cout});
        if (!st && m.find(prefix + n + "/bias")) st = check_shape(m, prefix + n + "/bias", {cout});
    }
    void bn(const std::string& n, int c) {
        for (const char* k : {"gamma", "beta", "mean", "var"}) if (!st) st = check_shape(m, prefix + n + "/" + k, {c});
    }
    void resnet(const std::string& n, int nd, int k, int c, bool use_bn) {
        for (int i = 0; i < 3; ++i) conv(n + "/conv" + std::to_string(i), nd, k, c, c);
        if (use_bn) { bn(n + "/bn0", c); bn(n + "/bn1", c); }
    }
    void dense(const std::string& n, int a, int b) {
        if (st) return;
        st = check_shape(m, prefix + n + "/kernel", {a, b});
        if (!st) st = check_shape(m, prefix + n + "/bias", {b});
    }
};

static int validate_weights(Model& m) {
    if (m.hp.present) {
        const HpnnCfg& h = m.hp;
        Expect e{m, "hpnn/", true};
        int cin = h.use_pos ? 3 : 1;
        for (size_t k = 0; k < h.pre.filters.size(); ++k) {
            e.conv("pre_bottleneck/" + std::to_string(k), 2, h.pre.ksizes[k], cin, h.pre.filters[k]);
            if (h.use_bn) e.bn("pre_bottleneck/" + std::to_string(k) + "/bn", h.pre.filters[k]);
            cin = h.pre.filters[k];
        }
        const int c0 = cin, F = h.F;
        for (const BlockCfg& b : h.blocks) {
            e.conv(b.name + "/conv0", 2, b.ksize, c0, F);
            for (int r = 1; r < b.n_convs; ++r) e.resnet(b.name + "/resnet" + std::to_string(r), 2, b.ksize, F, h.use_bn);
            if (b.deconv && !e.st) e.st = check_shape(m, "hpnn/" + b.name + "/deconv/kernel", {b.us, b.us, F, F});
        }
        e.conv("non_bottleneck_conv", 2, 5, c0, F);
        e.conv("post_merge_conv", 2, 7, 2 * F, F);
        e.resnet("post_merge_resnet", 2, 7, F, false);
        e.dense("dx_dense/0", 3, 100); e.dense("dx_dense/1", 100, 100); e.dense("dx_dense/2", 100, F);
        cin = F;
        const int S = (int)h.fin.filters.size();
        for (int k = 0; k < S; ++k) {
            e.conv("final/" + std::to_string(k) + "/conv", 2, h.fin.ksizes[k], cin, h.fin.filters[k]);
            if (k < S - h.fin.nreg) e.resnet("final/" + std::to_string(k) + "/resnet", 2, h.fin.ksizes[k], h.fin.filters[k], false);
            cin = h.fin.filters[k];
        }
        if (h.use_scaling) {
            cin = h.fin.filters.back() + 1;
            for (int s = 0; s < h.sc_stages; ++s) { e.conv("scaling/conv" + std::to_string(s), 2, h.sc_ksize, cin, h.sc_filters); cin = h.sc_filters; }
            const int nb = spp_bins(h.sc_levels, 2);
            e.dense("scaling/dense0", nb, 100); e.dense("scaling/dense1", 100, 25); e.dense("scaling/dense2", 25, 1);
        }
        if (e.st) return e.st;
    }
    if (m.db.present) {
        const DbcnnCfg& d = m.db;
        Expect e{m, "dbcnn/", true};
        int cin = 3;
        for (size_t k = 0; k < d.bnd.filters.size(); ++k) {
            const std::string n = "boundary/" + std::to_string(k);
            e.conv(n + "/conv", 1, d.bnd.ksizes[k], cin, d.bnd.filters[k]);
            if (d.use_bn) e.bn(n + "/bn", d.bnd.filters[k]);
            e.resnet(n + "/resnet", 1, d.bnd.ksizes[k], d.bnd.filters[k], d.use_bn);
            cin = d.bnd.filters[k];
        }
        int din = 3 + spp_bins(d.spp_levels, 1);
        for (size_t i = 0; i < d.mlp_units.size(); ++i) { e.dense("mlp/" + std::to_string(i), din, d.mlp_units[i]); din = d.mlp_units[i]; }
        cin = d.nmodes + 2;
        const int S = (int)d.fin.filters.size();
        for (int k = 0; k < S; ++k) {
            e.conv("final/" + std::to_string(k) + "/conv", 2, d.fin.ksizes[k], cin, d.fin.filters[k]);
            if (k < S - d.fin.nreg) e.resnet("final/" + std::to_string(k) + "/resnet", 2, d.fin.ksizes[k], d.fin.filters[k], false);
            cin = d.fin.filters[k];
        }
        if (e.st) return e.st;
    }
    return PCNN_OK;
}

static bool tc_kernel_ok(int k, float pad_value) { return k % 2 == 1 && k <= 15 && pad_value == 0.0f; }

static bool hpnn_tc_supported(const HpnnCfg& h) {
    for (int k : h.pre.ksizes) if (!tc_kernel_ok(k, 0.f)) return false;
    for (int k : h.fin.ksizes) if (!tc_kernel_ok(k, 0.f)) return false;
    for (const BlockCfg& b : h.blocks) if (b.deconv && !tc_kernel_ok(b.ksize, 0.f)) return false;
    int fmax = 0;
    for (int f : h.fin.filters) fmax = std::max(fmax, f);
    return h.F <= 32 && h.pre.pad_value == 0.f && h.fin.pad_value == 0.f && (h.pre.pad == 0 || h.pre.pad == 1) && h.fin.pad == 0 &&
           fmax <= 32 && h.pre.filters.size() >= 2;
}
static bool dbcnn_tc_supported(const DbcnnCfg& d) {
    for (int k : d.fin.ksizes) if (k % 2 == 0 || k > 15) return false;
    int fmax = 0;
    for (int f : d.fin.filters) fmax = std::max(fmax, f);
    return fmax <= 32 && d.nmodes + 2 <= 32 && d.fin.pad == PCNN_PAD_CONSTANT && d.fin.pad_value == 0.f;
}

// pack one conv layer for the tensor-core kernel (idempotent per (layer, nsplit))
static int tc_pack(Model& m, const std::string& name, int nsplit, cudaStream_t st) {
    const auto key = std::make_pair(name, nsplit);
    if (m.tc.count(key)) return PCNN_OK;
    ConvW c;
    TRY(m.conv(name, 2, &c));
    const size_t bytes = pcnn_conv_tc_packed_weight_bytes(c.k, c.k, c.cin, c.cout, nsplit);
    if (bytes == 0) { set_error("finalize_weights: layer '%s' (k %d, %d->%d) is not supported by the tensor-core kernel", name.c_str(), c.k, c.cin, c.cout); return PCNN_ERR_UNSUPPORTED; }
    TcPack p;
    TRY(m.dev_alloc(&p.packed, bytes));
    float amax = 0.f;
    for (float v : c.w->host) amax = std::max(amax, std::fabs(v));
    const float scale = pow2_prescale(amax);
    TRY(pcnn_conv_tc_pack_weights(c.kernel, p.packed, c.k, c.k, c.cin, c.cout, nsplit, scale, st));
    p.k = c.k; p.cin = c.cin; p.cout = c.cout; p.nsplit = nsplit; p.acc_scale = 1.0f / scale;
    m.tc[key] = p;
    return PCNN_OK;
}

static int finalize(Model& m, int precision, cudaStream_t st) {
    if (precision < PREC_FP32 || precision > PREC_MIXED) { set_error("finalize_weights: precision must be 0 (fp32), 1 (tc), 2 (tc3), 3 (tc2) or 4 (mixed)"); return PCNN_ERR_INVALID_ARGUMENT; }
    TRY(validate_weights(m));
    // fold BatchNorm: y = gamma (x - mean) / sqrt(var + 1e-3) + beta = x * scale + shift   (Keras epsilon 1e-3)
    if (m.bn.empty()) {
        for (auto& kv : m.w) {
            const std::string& name = kv.first;
            if (name.size() < 6 || name.compare(name.size() - 6, 6, "/gamma") != 0) continue;
            const std::string base = name.substr(0, name.size() - 6);
            const Weight *g = &kv.second, *b = m.find(base + "/beta"), *mu = m.find(base + "/mean"), *var = m.find(base + "/var");
            if (!b || !mu || !var) { set_error("finalize_weights: incomplete BatchNorm '%s'", base.c_str()); return PCNN_ERR_INVALID_ARGUMENT; }
            const size_t n = g->numel();
            std::vector<float> ss(2 * n);
            for (size_t i = 0; i < n; ++i) {
                const float scale = g->host[i] / std::sqrt(var->host[i] + 1e-3f);
                ss[i] = scale;
                ss[n + i] = b->host[i] - mu->host[i] * scale;
            }
            void* d = nullptr;
            TRY(m.dev_alloc(&d, 2 * n * sizeof(float)));
            PCNN_CHECK_CUDA(cudaMemcpy(d, ss.data(), 2 * n * sizeof(float), cudaMemcpyHostToDevice));
            m.bn[base] = {reinterpret_cast<float*>(d), reinterpret_cast<float*>(d) + n};
        }
    }
    const int mode_of[5] = {0, 1, 2, 3, 3};
    m.hp_mode = mode_of[precision];
    m.db_mode = precision == PREC_MIXED ? 1 : mode_of[precision];
    m.branch_single = precision == PREC_MIXED;
    if (precision != PREC_FP32) {
        if (m.hp.present) {
            const HpnnCfg& h = m.hp;
            if (!hpnn_tc_supported(h)) { set_error("finalize_weights: tensor-core precision covers odd kernels <= 15, <= 32 filters, zero CONSTANT / SYMMETRIC padding"); return PCNN_ERR_UNSUPPORTED; }
            const int nm = m.hp_mode, bm = m.branch_single ? 1 : nm;
            for (size_t k = 0; k < h.pre.filters.size(); ++k) TRY(tc_pack(m, "hpnn/pre_bottleneck/" + std::to_string(k), nm, st));
            for (const BlockCfg& b : h.blocks) {
                if (!tc_kernel_ok(b.ksize, b.pad_value) || b.pad > PCNN_PAD_SYMMETRIC) continue;      // FP32 kernels only
                TRY(tc_pack(m, "hpnn/" + b.name + "/conv0", bm, st));
                for (int r = 1; r < b.n_convs; ++r)
                    for (int i = 0; i < 3; ++i) TRY(tc_pack(m, "hpnn/" + b.name + "/resnet" + std::to_string(r) + "/conv" + std::to_string(i), bm, st));
            }
            TRY(tc_pack(m, "hpnn/non_bottleneck_conv", nm, st));
            TRY(tc_pack(m, "hpnn/post_merge_conv", nm, st));
            for (int i = 0; i < 3; ++i) TRY(tc_pack(m, "hpnn/post_merge_resnet/conv" + std::to_string(i), nm, st));
            const int S = (int)h.fin.filters.size();
            for (int k = 0; k < S; ++k) {
                TRY(tc_pack(m, "hpnn/final/" + std::to_string(k) + "/conv", nm, st));
                if (k < S - h.fin.nreg) for (int i = 0; i < 3; ++i) TRY(tc_pack(m, "hpnn/final/" + std::to_string(k) + "/resnet/conv" + std::to_string(i), nm, st));
            }
            // packed transpose-conv kernels of the fused upsample-merge
            for (const BlockCfg& b : h.blocks) {
                if (!b.deconv) continue;
                const std::string n = "hpnn/" + b.name + "/deconv";
                const Weight* k = m.find(n + "/kernel");
                if (!m.deconv_tc.count(n) && h.F == 32 && b.us <= 32 && pcnn_upsample_merge_tc_packed_bytes(b.us)) {
                    void* p = nullptr;
                    TRY(m.dev_alloc(&p, pcnn_upsample_merge_tc_packed_bytes(b.us)));
                    TRY(pcnn_upsample_merge_tc_pack_kernel(k->dev, p, b.us, st));
                    m.deconv_tc[n] = p;
                }
                if (!m.deconv_f32.count(n) && h.F % 8 == 0 && b.us <= 32 && pcnn_upsample_merge_packed_floats(b.us, h.F)) {
                    void* p = nullptr;
                    TRY(m.dev_alloc(&p, pcnn_upsample_merge_packed_floats(b.us, h.F) * sizeof(float)));
                    TRY(pcnn_upsample_merge_pack_kernel(k->dev, reinterpret_cast<float*>(p), b.us, h.F, st));
                    m.deconv_f32[n] = p;
                }
            }
        }
        if (m.db.present) {
            const DbcnnCfg& d = m.db;
            if (!dbcnn_tc_supported(d)) { set_error("finalize_weights: tensor-core precision covers odd kernels <= 15, <= 32 filters, zero CONSTANT padding"); return PCNN_ERR_UNSUPPORTED; }
            const int S = (int)d.fin.filters.size();
            for (int k = 0; k < S; ++k) {
                TRY(tc_pack(m, "dbcnn/final/" + std::to_string(k) + "/conv", m.db_mode, st));
                if (k < S - d.fin.nreg) for (int i = 0; i < 3; ++i) TRY(tc_pack(m, "dbcnn/final/" + std::to_string(k) + "/resnet/conv" + std::to_string(i), m.db_mode, st));
            }
        }
    }
    PCNN_CHECK_CUDA(cudaStreamSynchronize(st));
    m.precision = precision;
    m.finalized = true;
    m.prepared.clear();          // tensors of another precision mode have different slot classes and tables
    m.need_cache.clear();
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------------ tables in the workspace
// A table is built on the host once per (workspace, shape) and lives in a never-released slot of the arena.
template <typename T, typename Fn>
static int table(Ctx& c, const std::string& name, size_t count, Fn build, const T** out) {
    auto it = c.prep->tables.find(name);
    int slot;
    if (it == c.prep->tables.end()) {
        slot = c.alloc(SlotKey{3, (int)c.prep->tables.size(), 0, 0, 0, 0, 0}, count * sizeof(T));
        c.prep->tables[name] = slot;
        if (!c.dry) {
            const std::vector<T> host = build();
            if (host.size() != count) { set_error("internal: table '%s' has %zu entries, expected %zu", name.c_str(), host.size(), count); return PCNN_ERR_INVALID_ARGUMENT; }
            // pageable source: the runtime has staged the bytes when the call returns (host-blocking, first call of a shape only)
            PCNN_CHECK_CUDA(cudaMemcpyAsync(c.ptr(slot), host.data(), count * sizeof(T), cudaMemcpyHostToDevice, c.st));
        }
    } else {
        slot = it->second;
    }
    *out = reinterpret_cast<const T*>(c.ptr(slot));
    return PCNN_OK;
}

static int pos_table(Ctx& c, int n, const float** out) {
    return table<float>(c, "pos/" + std::to_string(n), (size_t)n, [n] { return position_table(n); }, out);
}

struct DevAxis { const int32_t* idx = nullptr; const float* w = nullptr; int taps = 0; };
static int axis_table(Ctx& c, int n_in, int n_out, int method, DevAxis* out) {
    const int taps = method == RESIZE_NEAREST ? 1 : (method == RESIZE_BILINEAR ? 2 : 4);
    const std::string key = "rs/" + std::to_string(n_in) + "/" + std::to_string(n_out) + "/" + std::to_string(method);
    out->taps = taps;
    TRY((table<int32_t>(c, key + "/i", (size_t)n_out * taps, [=] { return resize_axis_table(n_in, n_out, method).idx; }, &out->idx)));
    TRY((table<float>(c, key + "/w", (size_t)n_out * taps, [=] { return resize_axis_table(n_in, n_out, method).w; }, &out->w)));
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------------ op wrappers
static int conv_f32(Ctx& c, const F32& x, const ConvW& w, int act, int pad, float pad_value, const F32* residual,
                    const float* out_scale, F32* out, int Bcap, F32* dst = nullptr) {
    if (w.cin != x.C) { set_error("conv2d: kernel expects %d input channels, got %d", w.cin, x.C); return PCNN_ERR_INVALID_ARGUMENT; }
    F32 y = dst ? *dst : c.f32(Bcap, x.B, w.cout, x.H, x.W);
    const int kh = x.H == 1 && w.w->shape.size() == 3 ? 1 : w.k;      // Conv1D runs as the H == 1, kh == 1 case
    RUN(c, pcnn_conv2d_f32(x.p, w.kernel, w.bias, w.bn_scale, w.bn_shift, residual ? residual->p : nullptr, out_scale, y.p, x.B,
                           x.C, w.cout, x.H, x.W, kh, w.k, pad, pad_value, act, x.bs, y.bs, residual ? residual->bs : 0, c.st));
    *out = y;
    return PCNN_OK;
}

// blocks/resnet.py:29-39 with BN and the residual add fused into the conv epilogues; consumes x
static int resnet_f32(Ctx& c, F32& x, const std::string& name, int nd, int act, int pad, float pad_value, bool use_bn,
                      const float* out_scale, int Bcap, F32* out) {
    ConvW w0, w1, w2;
    TRY(c.m->conv(name + "/conv0", nd, &w0)); TRY(c.m->conv(name + "/conv1", nd, &w1)); TRY(c.m->conv(name + "/conv2", nd, &w2));
    if (use_bn) { c.m->with_bn(name + "/bn0", &w0); c.m->with_bn(name + "/bn1", &w1); }
    F32 t, u;
    TRY(conv_f32(c, x, w0, act, pad, pad_value, nullptr, nullptr, &t, Bcap));
    TRY(conv_f32(c, t, w1, act, pad, pad_value, &x, nullptr, &u, Bcap));
    c.free(t); c.free(x);
    TRY(conv_f32(c, u, w2, act, pad, pad_value, nullptr, out_scale, out, Bcap));
    c.free(u);
    return PCNN_OK;
}

static int avgpool(Ctx& c, const F32& x, int s, int Bcap, F32* out) {
    F32 y = c.f32(Bcap, x.B, x.C, cdiv(x.H, s), cdiv(x.W, s));
    RUN(c, pcnn_avgpool_same_f32(x.p, y.p, x.B, x.C, x.H, x.W, s, x.bs, c.st));
    *out = y;
    return PCNN_OK;
}

static int dense(Ctx& c, const float* x, int B, int Bcap, const std::string& name, int act, float** y, int* nout) {
    const Weight* k = c.m->find(name + "/kernel");
    const Weight* b = c.m->find(name + "/bias");
    if (!k) { set_error("weight '%s/kernel' was never set", name.c_str()); return PCNN_ERR_INVALID_ARGUMENT; }
    float* out = c.vec((size_t)Bcap * k->shape[1]);
    RUN(c, pcnn_dense_f32(x, k->dev, b ? b->dev : nullptr, out, B, (int)k->shape[0], (int)k->shape[1], act, c.st));
    *y = out;
    *nout = (int)k->shape[1];
    return PCNN_OK;
}

static int halo_fill(Ctx& c, B8& t, int pad, int mode) {
    if (mode == PCNN_PAD_CONSTANT && t.halo.mode == PCNN_PAD_CONSTANT) return PCNN_OK;     // zero halo is 7 wide from the start
    if (t.halo.mode == mode && t.halo.pad >= pad) return PCNN_OK;                          // a wider mirrored ring contains it
    if (mode == PCNN_PAD_CONSTANT) pad = 7;
    RUN(c, pcnn_blk8_halo_fill(t.hi, t.B, t.C, t.H, t.W, pad, mode, c.st));
    if (t.s_lo >= 0) RUN(c, pcnn_blk8_halo_fill(t.lo, t.B, t.C, t.H, t.W, pad, mode, c.st));
    t.halo.mode = mode; t.halo.pad = pad;
    return PCNN_OK;
}

// NCHW fp32 -> BLK8 (new tensor, or channels [c_offset, c_offset+C) of `into`)
static int to_blk8(Ctx& c, const F32& x, int Bcap, int mode, int halo, B8* into, int c_offset, B8* out) {
    B8 t = into ? *into : c.b8(Bcap, x.B, x.C, x.H, x.W, mode, fusable_halo(halo, x.H, x.W));
    const bool fused = fusable_halo(halo, x.H, x.W) && (t.C + 15) / 16 == (x.C + 15) / 16 && c_offset == 0;
    RUN(c, pcnn_to_blk8(x.p, t.hi, t.lo, t.mode, x.B, x.C, x.H, x.W, t.C, c_offset, x.bs, fused ? PCNN_PAD_SYMMETRIC : PCNN_PAD_CONSTANT, c.st));
    if (fused) { t.halo.mode = PCNN_PAD_SYMMETRIC; t.halo.pad = 7; }
    else if (t.halo.mode != PCNN_PAD_CONSTANT) t.halo.pad = -1;
    if (into) *into = t;
    if (out) *out = t;
    return PCNN_OK;
}

static int from_blk8(Ctx& c, const B8& t, int C, int Bcap, F32* dst, F32* out) {
    F32 y = dst ? *dst : c.f32(Bcap, t.B, C, t.H, t.W);
    RUN(c, pcnn_from_blk8(t.hi, t.lo, t.mode, y.p, t.B, C, t.H, t.W, t.C, 0, y.bs, c.st));
    *out = y;
    return PCNN_OK;
}

struct TcArgs {
    int act = 0, pad = PCNN_PAD_CONSTANT, next_pad = PCNN_PAD_CONSTANT;
    const char* bn = nullptr;          // name of the BatchNorm that follows, or null
    const B8* residual = nullptr;
    const float* out_scale = nullptr;
    B8* into = nullptr;                // write channels [0, cout) of this tensor (in-place concat)
};

static int conv_tc(Ctx& c, B8& x, const std::string& name, const TcArgs& a, int Bcap, B8* out) {
    Model& m = *c.m;
    auto it = m.tc.find(std::make_pair(name, x.mode));
    if (it == m.tc.end()) { set_error("internal: layer '%s' was not packed for precision mode %d", name.c_str(), x.mode); return PCNN_ERR_INVALID_ARGUMENT; }
    const TcPack& wp = it->second;
    if (cdiv(wp.cin, 16) != cdiv(x.C, 16)) { set_error("conv2d_tc: kernel '%s' expects %d input channels, tensor holds %d", name.c_str(), wp.cin, x.C); return PCNN_ERR_INVALID_ARGUMENT; }
    const Weight* bias = m.find(name + "/bias");
    const float *bn_s = nullptr, *bn_t = nullptr;
    if (a.bn) { auto b = m.bn.find(a.bn); if (b != m.bn.end()) { bn_s = b->second.first; bn_t = b->second.second; } }
    TRY(halo_fill(c, x, wp.k / 2, a.pad));
    B8 y = a.into ? *a.into : c.b8(Bcap, x.B, wp.cout, x.H, x.W, x.mode, fusable_halo(a.next_pad, x.H, x.W));
    const bool split = x.mode >= 2;
    const bool fused_halo = fusable_halo(a.next_pad, x.H, x.W) && (y.C + 15) / 16 == (wp.cout + 15) / 16;
    const bool timed = !c.dry && m.prof_k == wp.k && m.prof_cin == wp.cin && m.prof_cout == wp.cout && m.prof_used + 2 <= m.prof_ev.size();
    if (timed) PCNN_CHECK_CUDA(cudaEventRecord(m.prof_ev[m.prof_used], c.st));
    RUN(c, pcnn_conv2d_tc(x.hi, split ? x.lo : nullptr, wp.packed, bias ? bias->dev : nullptr, bn_s, bn_t,
                          a.residual ? a.residual->hi : nullptr, (a.residual && split) ? a.residual->lo : nullptr, a.out_scale,
                          y.hi, split ? y.lo : nullptr, x.B, wp.cin, wp.cout, y.C, a.residual ? a.residual->C : 0, x.H, x.W, wp.k,
                          a.act, x.mode, wp.acc_scale, fused_halo ? PCNN_PAD_SYMMETRIC : PCNN_PAD_CONSTANT, c.num_sms, c.st));
    if (timed) {
        PCNN_CHECK_CUDA(cudaEventRecord(m.prof_ev[m.prof_used + 1], c.st));
        m.prof_used += 2;
        m.prof_flops += 2.0 * x.B * x.H * x.W * wp.k * wp.k * wp.cin * wp.cout;
    }
    if (fused_halo) { y.halo.mode = PCNN_PAD_SYMMETRIC; y.halo.pad = 7; }
    else if (y.halo.mode != PCNN_PAD_CONSTANT) y.halo.pad = -1;
    if (a.into) *a.into = y;
    *out = y;
    return PCNN_OK;
}

// resnet on BLK8 tensors; consumes x
static int resnet_tc(Ctx& c, B8& x, const std::string& name, int act, int pad, bool use_bn, const float* out_scale,
                     int next_pad, int Bcap, B8* out) {
    const std::string b0 = name + "/bn0", b1 = name + "/bn1";
    TcArgs a0; a0.act = act; a0.pad = pad; a0.next_pad = pad; a0.bn = use_bn ? b0.c_str() : nullptr;
    B8 t, u;
    TRY(conv_tc(c, x, name + "/conv0", a0, Bcap, &t));
    TcArgs a1 = a0; a1.bn = use_bn ? b1.c_str() : nullptr; a1.residual = &x;
    TRY(conv_tc(c, t, name + "/conv1", a1, Bcap, &u));
    c.free(t); c.free(x);
    TcArgs a2; a2.act = act; a2.pad = pad; a2.next_pad = next_pad; a2.out_scale = out_scale;
    TRY(conv_tc(c, u, name + "/conv2", a2, Bcap, out));
    c.free(u);
    return PCNN_OK;
}

static int stack_layers(const Model& m, const std::string& first_conv, const std::vector<std::string>& resnets, int nd,
                        const char* first_bn, bool use_bn, StackLayers* L) {
    auto push = [&](const std::string& conv, const std::string& bn, bool with_bn, int flags) -> int {
        ConvW w;
        TRY(m.conv(conv, nd, &w));
        if (with_bn) m.with_bn(bn, &w);
        L->kernels.push_back(w.kernel); L->biases.push_back(w.bias); L->bn_scale.push_back(w.bn_scale); L->bn_shift.push_back(w.bn_shift);
        L->ksize.push_back(w.k); L->cin.push_back(w.cin); L->cout.push_back(w.cout); L->flags.push_back(flags);
        return PCNN_OK;
    };
    TRY(push(first_conv, first_bn ? first_bn : "", first_bn != nullptr, 0));
    const int fl[3] = {1, 2, 0};
    for (const std::string& rn : resnets)
        for (int i = 0; i < 3; ++i) TRY(push(rn + "/conv" + std::to_string(i), rn + "/bn" + std::to_string(i), use_bn && i < 2, fl[i]));
    return PCNN_OK;
}

static bool smallmap_supported(int H, int W, const StackLayers& L) {
    if (L.n() == 0 || L.n() > 24 || H * W > 64) return false;
    int pm = 0;
    size_t wmax = 0;
    for (int i = 0; i < L.n(); ++i) {
        if (L.ksize[i] % 2 == 0 || L.ksize[i] > 7 || std::max(L.cin[i], L.cout[i]) > 32) return false;
        pm = std::max(pm, L.ksize[i] / 2);
        wmax = std::max(wmax, (size_t)L.ksize[i] * L.ksize[i] * L.cin[i] * 32);
    }
    return ((size_t)3 * 32 * (H + 2 * pm) * (W + 2 * pm) + wmax) * 4 <= 220 * 1024;
}

static bool boundary_supported(int n, const StackLayers& L) {
    if (L.n() == 0 || L.n() > 48) return false;
    size_t wmax = 0;
    for (int i = 0; i < L.n(); ++i) {
        if (L.ksize[i] % 2 == 0 || L.ksize[i] > 19 || std::max(L.cin[i], L.cout[i]) > 28) return false;
        wmax = std::max(wmax, (size_t)L.ksize[i] * L.cin[i] * 32);
    }
    return ((size_t)3 * 28 * ((n + 18 + 3) / 4 * 4) + wmax) * 4 <= 220 * 1024;
}

static int jacobi(Ctx& c, float* cur_in, const float* rhs, const float* dx, int B, int Bcap, int H, int W, int iters, float* final_out) {
    // grid_spacings [B,2] = (dx, dx); ping-pong; the last sweep lands in final_out
    float* gs = c.vec((size_t)Bcap * 2);
    if (!c.dry) {
        PCNN_CHECK_CUDA(cudaMemcpy2DAsync(gs, 8, dx, 4, 4, B, cudaMemcpyDeviceToDevice, c.st));
        PCNN_CHECK_CUDA(cudaMemcpy2DAsync(gs + 1, 8, dx, 4, 4, B, cudaMemcpyDeviceToDevice, c.st));
    }
    float* tmp = iters > 1 ? c.vec((size_t)Bcap * H * W) : nullptr;
    const float* src = cur_in;
    for (int i = 0; i < iters; ++i) {
        float* dst = ((iters - 1 - i) % 2 == 0) ? final_out : tmp;
        RUN(c, pcnn_jacobi_sweep_f32(src, rhs, gs, dst, B, H, W, c.st));
        src = dst;
    }
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------------ HPNN
static int out_size(int n, int ds, int us) { return (int)(((double)n / (double)ds) * (double)us); }    // bottleneck_block.py:82

static int check_branch_shapes(const HpnnCfg& h, int H, int W) {
    for (const BlockCfg& b : h.blocks) {
        const int oh = out_size(H, b.ds, b.us), ow = out_size(W, b.ds, b.us);
        if (oh != H || ow != W) {
            set_error("bottleneck branch ds=%d would produce (%d, %d) for a (%d, %d) grid; the merge needs equal shapes", b.ds, oh, ow, H, W);
            return PCNN_ERR_INVALID_ARGUMENT;
        }
    }
    return PCNN_OK;
}

// pool -> conv -> resnets of one branch in FP32 (blocks/bottleneck_block.py:36-50); consumes nothing (pooled stays alive)
static int branch_stack_layers(Ctx& c, const BlockCfg& b, std::vector<std::string>* rns, StackLayers* L) {
    const std::string name = "hpnn/" + b.name;
    rns->clear();
    for (int r = 1; r < b.n_convs; ++r) rns->push_back(name + "/resnet" + std::to_string(r));
    return stack_layers(*c.m, name + "/conv0", *rns, 2, nullptr, b.use_bn, L);
}

static int branch_lowres_f32(Ctx& c, const BlockCfg& b, const F32& pooled, int Bcap, F32* out) {
    const std::string name = "hpnn/" + b.name;
    std::vector<std::string> rns;
    StackLayers L;
    TRY(branch_stack_layers(c, b, &rns, &L));
    if (smallmap_supported(pooled.H, pooled.W, L) && pooled.bs == (long long)pooled.C * pooled.H * pooled.W) {
        F32 y = c.f32(Bcap, pooled.B, L.cout.back(), pooled.H, pooled.W);
        RUN(c, pcnn_smallmap_stack_f32(pooled.p, y.p, pooled.B, pooled.H, pooled.W, pooled.C, L.n(), L.kernels.data(), L.biases.data(),
                                       L.bn_scale.data(), L.bn_shift.data(), L.ksize.data(), L.cin.data(), L.cout.data(), L.flags.data(),
                                       b.act, b.pad, b.pad_value, c.st));
        *out = y;
        return PCNN_OK;
    }
    ConvW w;
    TRY(c.m->conv(name + "/conv0", 2, &w));
    F32 hcur;
    TRY(conv_f32(c, pooled, w, b.act, b.pad, b.pad_value, nullptr, nullptr, &hcur, Bcap));
    for (const std::string& rn : rns) {
        F32 nxt;
        TRY(resnet_f32(c, hcur, rn, 2, b.act, b.pad, b.pad_value, b.use_bn, nullptr, Bcap, &nxt));
        hcur = nxt;
    }
    *out = hcur;
    return PCNN_OK;
}

// the last linear convs (strict mode), Scaling, boundary ring and post-smoother; consumes y
static int hpnn_tail(Ctx& c, F32 y, bool y_in_cat2, F32 cat2, const float* rhs, const float* dx, int first_regular, int B,
                     int Bcap, int H, int W, float* out) {
    const HpnnCfg& h = c.m->hp;
    const int S = (int)h.fin.filters.size();
    if (cat2.slot < 0 && h.use_scaling) cat2 = c.f32(Bcap, B, 2, H, W);
    for (int k = first_regular; k < S; ++k) {
        ConvW w;
        TRY(c.m->conv("hpnn/final/" + std::to_string(k) + "/conv", 2, &w));
        const bool last = (k == S - 1) && h.use_scaling && w.cout == 1;
        F32 view = cat2; view.C = 1; view.slot = -1;
        F32 nxt;
        TRY(conv_f32(c, y, w, PCNN_ACT_LINEAR, PCNN_PAD_CONSTANT, 0.f, nullptr, nullptr, &nxt, Bcap, last ? &view : nullptr));
        if (!y_in_cat2) c.free(y);
        y = nxt;
        y_in_cat2 = last;
    }
    const float* sdev = nullptr;
    if (h.use_scaling) {
        if (!y_in_cat2) {
            if (y.C != 1) { set_error("Scaling needs a single-channel network output, got %d channels", y.C); return PCNN_ERR_INVALID_ARGUMENT; }
            if (!c.dry) PCNN_CHECK_CUDA(cudaMemcpy2DAsync(cat2.p, (size_t)2 * H * W * 4, y.p, (size_t)y.bs * 4, (size_t)H * W * 4, B, cudaMemcpyDeviceToDevice, c.st));
            c.free(y);
            y = cat2; y.C = 1; y.slot = -1;
            y_in_cat2 = true;
        }
        if (!c.dry) PCNN_CHECK_CUDA(cudaMemcpy2DAsync(cat2.p + (size_t)H * W, (size_t)2 * H * W * 4, rhs, (size_t)H * W * 4, (size_t)H * W * 4, B, cudaMemcpyDeviceToDevice, c.st));
        F32 hcur = cat2;
        hcur.slot = -1;                       // cat2 stays alive until the finalize kernel has read channel 0
        for (int st = 0; st < h.sc_stages; ++st) {
            ConvW w;
            TRY(c.m->conv("hpnn/scaling/conv" + std::to_string(st), 2, &w));
            F32 t, p;
            TRY(conv_f32(c, hcur, w, h.sc_act, PCNN_PAD_CONSTANT, 0.f, nullptr, nullptr, &t, Bcap));
            c.free(hcur);
            TRY(avgpool(c, t, h.sc_ratio, Bcap, &p));
            c.free(t);
            hcur = p;
        }
        const int nb = spp_bins(h.sc_levels, 2);
        const int32_t* boxes = nullptr;
        const int hh = hcur.H, ww = hcur.W;
        const std::vector<std::vector<int>>& lv = h.sc_levels;
        TRY((table<int32_t>(c, "sppbox2/" + std::to_string(hh) + "/" + std::to_string(ww), (size_t)nb * 4, [&lv, hh, ww] { return spp_boxes(lv, hh, ww, 2); }, &boxes)));
        float* v = c.vec((size_t)Bcap * nb);
        RUN(c, pcnn_spp_f32(hcur.p, boxes, v, B, hcur.C, hcur.H, hcur.W, nb, PCNN_POOL_MAX, c.st));
        c.free(hcur);
        float *v1, *v2, *v3;
        int n;
        TRY(dense(c, v, B, Bcap, "hpnn/scaling/dense0", PCNN_ACT_LEAKY_RELU, &v1, &n));
        TRY(dense(c, v1, B, Bcap, "hpnn/scaling/dense1", PCNN_ACT_LEAKY_RELU, &v2, &n));
        TRY(dense(c, v2, B, Bcap, "hpnn/scaling/dense2", PCNN_ACT_LINEAR, &v3, &n));
        sdev = v3;
    }
    if (h.postsmooth > 0) {
        float* pre = c.vec((size_t)Bcap * H * W);
        RUN(c, pcnn_hpnn_finalize_f32(y.p, sdev, pre, B, H, W, h.bc_type, y.bs, c.st));
        TRY(jacobi(c, pre, rhs, dx, B, Bcap, H, W, h.postsmooth, out));
    } else {
        RUN(c, pcnn_hpnn_finalize_f32(y.p, sdev, out, B, H, W, h.bc_type, y.bs, c.st));
    }
    if (!y_in_cat2) c.free(y);
    c.free(cat2);
    return PCNN_OK;
}

static int dx_mlp(Ctx& c, const float* dx, int B, int Bcap, int H, int W, float** d) {
    float* in = c.vec((size_t)Bcap * 3);
    RUN(c, pcnn_dense_input_f32(dx, nullptr, in, B, H, W, 0, 0, c.st));
    float *a, *b;
    int n;
    TRY(dense(c, in, B, Bcap, "hpnn/dx_dense/0", PCNN_ACT_LEAKY_RELU, &a, &n));
    TRY(dense(c, a, B, Bcap, "hpnn/dx_dense/1", PCNN_ACT_LEAKY_RELU, &b, &n));
    TRY(dense(c, b, B, Bcap, "hpnn/dx_dense/2", PCNN_ACT_LINEAR, d, &n));
    return PCNN_OK;
}

static int hpnn_input(Ctx& c, const float* rhs, int B, int Bcap, int H, int W, F32* x) {
    const HpnnCfg& h = c.m->hp;
    if (!h.use_pos) {
        x->slot = -1; x->p = const_cast<float*>(rhs); x->B = B; x->C = 1; x->H = H; x->W = W; x->bs = (long long)H * W;
        return PCNN_OK;
    }
    const float *px, *py;
    TRY(pos_table(c, H, &px));
    TRY(pos_table(c, W, &py));
    *x = c.f32(Bcap, B, 3, H, W);
    RUN(c, pcnn_hpnn_input_f32(rhs, px, py, x->p, B, H, W, c.st));
    return PCNN_OK;
}

// strict FP32 program (Homogeneous_Poisson_NN_Legacy.py:182-257)
static int hpnn_fp32(Ctx& c, const float* rhs, const float* dx, float* out, int B, int Bcap, int H, int W) {
    const HpnnCfg& h = c.m->hp;
    const Model& m = *c.m;
    TRY(check_branch_shapes(h, H, W));
    F32 x;
    TRY(hpnn_input(c, rhs, B, Bcap, H, W, &x));
    for (size_t k = 0; k < h.pre.filters.size(); ++k) {
        const std::string n = "hpnn/pre_bottleneck/" + std::to_string(k);
        ConvW w;
        TRY(m.conv(n, 2, &w));
        if (h.use_bn) m.with_bn(n + "/bn", &w);
        F32 y;
        TRY(conv_f32(c, x, w, h.pre.act, h.pre.pad, h.pre.pad_value, nullptr, nullptr, &y, Bcap));
        c.free(x);
        x = y;
    }
    F32 x0 = x;
    const int F = h.F;
    // concat(non_bottleneck_conv(x0), merged) is assembled in place: channels [0,F) and [F,2F)
    F32 cat = c.f32(Bcap, B, 2 * F, H, W);
    F32 merged = cat; merged.slot = -1; merged.p = cat.p ? cat.p + (size_t)F * H * W : nullptr; merged.C = F;
    const float alpha = 1.0f / (float)(h.blocks.size() * F);
    bool first = true;
    for (const BlockCfg& b : h.blocks) {
        F32 pooled, low;
        TRY(avgpool(c, x0, b.ds, Bcap, &pooled));
        TRY(branch_lowres_f32(c, b, pooled, Bcap, &low));
        c.free(pooled);
        if (b.deconv) {
            const Weight* k = m.find("hpnn/" + b.name + "/deconv/kernel");
            const Weight* bias = m.find("hpnn/" + b.name + "/deconv/bias");
            RUN(c, pcnn_deconv_same_f32(low.p, k->dev, bias ? bias->dev : nullptr, merged.p, B, F, F, low.H, low.W, H, W, b.us, b.us,
                                        b.us, b.deconv_act, alpha, first ? 0 : 1, merged.bs, c.st));
        } else {
            DevAxis ay, ax;
            TRY(axis_table(c, low.H, H, b.resize_method, &ay));
            TRY(axis_table(c, low.W, W, b.resize_method, &ax));
            RUN(c, pcnn_resize_f32(low.p, ay.idx, ay.w, ax.idx, ax.w, ay.taps, merged.p, B, F, low.H, low.W, H, W, alpha, first ? 0 : 1, merged.bs, c.st));
        }
        c.free(low);
        first = false;
    }
    {
        ConvW w;
        TRY(m.conv("hpnn/non_bottleneck_conv", 2, &w));
        F32 view = cat; view.slot = -1; view.C = F;
        F32 dummy;
        TRY(conv_f32(c, x0, w, PCNN_ACT_LEAKY_RELU, PCNN_PAD_CONSTANT, 0.f, nullptr, nullptr, &dummy, Bcap, &view));
        c.free(x0);
    }
    F32 y;
    {
        ConvW w;
        TRY(m.conv("hpnn/post_merge_conv", 2, &w));
        TRY(conv_f32(c, cat, w, PCNN_ACT_LEAKY_RELU, PCNN_PAD_CONSTANT, 0.f, nullptr, nullptr, &y, Bcap));
        c.free(cat);
    }
    float* d;
    TRY(dx_mlp(c, dx, B, Bcap, H, W, &d));
    {
        F32 nxt;
        TRY(resnet_f32(c, y, "hpnn/post_merge_resnet", 2, PCNN_ACT_LEAKY_RELU, PCNN_PAD_CONSTANT, 0.f, false, d, Bcap, &nxt));
        y = nxt;
    }
    const int S = (int)h.fin.filters.size(), nreg = h.fin.nreg;
    for (int k = 0; k < S - nreg; ++k) {
        const std::string n = "hpnn/final/" + std::to_string(k);
        ConvW w;
        TRY(m.conv(n + "/conv", 2, &w));
        F32 t, nxt;
        TRY(conv_f32(c, y, w, h.fin.act, h.fin.pad, h.fin.pad_value, nullptr, nullptr, &t, Bcap));
        c.free(y);
        TRY(resnet_f32(c, t, n + "/resnet", 2, h.fin.act, PCNN_PAD_CONSTANT, 0.f, false, nullptr, Bcap, &nxt));
        y = nxt;
    }
    F32 none;
    return hpnn_tail(c, y, false, none, rhs, dx, S - nreg, B, Bcap, H, W, out);
}

// {s: AveragePooling2D(s, 'same')(x0)}; where the windows nest exactly (grid divisible by s) level s is pooled from the
// largest already computed level that divides it, so x0 is read once or twice instead of once per branch
static int pool_pyramid(Ctx& c, const F32& x0, const std::vector<int>& factors, int Bcap, std::map<int, F32>* levels) {
    std::vector<int> fs = factors;
    std::sort(fs.begin(), fs.end());
    fs.erase(std::unique(fs.begin(), fs.end()), fs.end());
    for (int s : fs) {
        const F32* src = &x0;
        int f = s;
        if (x0.H % s == 0 && x0.W % s == 0) {
            for (auto it = levels->rbegin(); it != levels->rend(); ++it) {
                const int t = it->first;
                if (s % t == 0 && x0.H % t == 0 && x0.W % t == 0) { src = &it->second; f = s / t; break; }
            }
        }
        F32 p;
        TRY(avgpool(c, *src, f, Bcap, &p));
        (*levels)[s] = p;
    }
    return PCNN_OK;
}

// tensor-core program: every heavy convolution on tcgen05 (BLK8 fp16 activations)
static int hpnn_tc(Ctx& c, const float* rhs, const float* dx, float* out, int B, int Bcap, int H, int W) {
    Model& m = *c.m;
    const HpnnCfg& h = m.hp;
    TRY(check_branch_shapes(h, H, W));
    const int F = h.F, split = m.hp_mode;
    const int npre = (int)h.pre.filters.size();
    F32 x;
    TRY(hpnn_input(c, rhs, B, Bcap, H, W, &x));
    B8 t;
    TRY(to_blk8(c, x, Bcap, split, h.pre.pad, nullptr, 0, &t));
    c.free(x);
    for (int k = 0; k < npre; ++k) {
        const std::string n = "hpnn/pre_bottleneck/" + std::to_string(k), bn = n + "/bn";
        TcArgs a; a.act = h.pre.act; a.pad = h.pre.pad; a.bn = h.use_bn ? bn.c_str() : nullptr;
        a.next_pad = k + 1 < npre ? h.pre.pad : PCNN_PAD_CONSTANT;
        B8 y;
        TRY(conv_tc(c, t, n, a, Bcap, &y));
        c.free(t);
        t = y;
    }
    B8 x0 = t;
    F32 x0f;
    TRY(from_blk8(c, x0, x0.C, Bcap, nullptr, &x0f));      // the pooling pyramid reads NCHW fp32
    std::vector<int> factors;
    for (const BlockCfg& b : h.blocks) factors.push_back(b.ds);
    std::map<int, F32> pools;
    TRY(pool_pyramid(c, x0f, factors, Bcap, &pools));
    c.free(x0f);

    const float alpha = 1.0f / (float)(h.blocks.size() * F);
    // 'mixed': the branches are averaged with weight 1/(8F) before they re-enter the trunk and run single-pass
    const int bsplit = m.branch_single ? 1 : split;
    int ndeconv = 0;
    bool strides_ok = true;
    std::vector<int> um_strides, um_ih, um_iw;
    for (const BlockCfg& b : h.blocks) {
        if (b.deconv) { ++ndeconv; strides_ok = strides_ok && b.us <= 32; um_strides.push_back(b.us); }
        else { um_ih.push_back(cdiv(H, b.ds)); um_iw.push_back(cdiv(W, b.ds)); }
    }
    // um_tc 1: every branch in the fused tensor-core upsample-merge (its resize sources are staged whole in shared memory:
    // grids up to ~400 pixels a side); 2: larger grids -- transpose-conv branches fused, resize branches added by a second
    // pass (pcnn_resize_add_blk8); 0: general kernels
    int um_tc = (bsplit == 1 && F == 32 && ndeconv > 0 && h.blocks.size() <= 16 && strides_ok) ? 1 : 0;
    if (um_tc) {
        auto fits = [&](int nrs) {
            const size_t n = pcnn_upsample_merge_tc_smem_bytes((int)um_strides.size(), um_strides.data(), nrs, um_ih.data(), um_iw.data());
            return n > 0 && n <= 227 * 1024;
        };
        um_tc = fits((int)um_ih.size()) ? 1 : ((!um_ih.empty() && fits(0)) ? 2 : 0);
    }
    struct Branch { const BlockCfg* cfg; bool is_b8; B8 b8; F32 f32; };
    std::vector<Branch> br;
    // tensor cores for every branch whose pooled map is at least 16 pixels a side (Python: _branch_on_tc)
    auto branch_on_tc = [&](const BlockCfg& b) {
        return std::min(cdiv(H, b.ds), cdiv(W, b.ds)) >= 16 && tc_kernel_ok(b.ksize, b.pad_value) && b.pad <= PCNN_PAD_SYMMETRIC;
    };
    // The small-map branches (pooled maps of at most 64 pixels: the FP32 stack kernel, one CTA per sample, ~0.35 ms of latency
    // each) all go into ONE launch, side by side: 3 branches at 256 x 256, 5 at 64 x 64.  Same arithmetic as one launch per
    // branch (the op-by-op Python program), so the results stay bit-identical.
    std::map<const BlockCfg*, F32> small_out;
    {
        std::vector<StackLayers> progs;
        std::vector<sms::StackDesc> descs;
        progs.reserve(h.blocks.size());
        std::vector<std::string> rns;
        for (const BlockCfg& b : h.blocks) {
            if (branch_on_tc(b) || (int)descs.size() == sms::MAX_PROGRAMS) continue;
            const F32& pooled = pools[b.ds];
            StackLayers L;
            TRY(branch_stack_layers(c, b, &rns, &L));
            if (!smallmap_supported(pooled.H, pooled.W, L) || pooled.bs != (long long)pooled.C * pooled.H * pooled.W) continue;
            progs.push_back(L);
            const StackLayers& S = progs.back();
            F32 y = c.f32(Bcap, pooled.B, S.cout.back(), pooled.H, pooled.W);
            descs.push_back(sms::StackDesc{pooled.p, y.p, pooled.H, pooled.W, pooled.C, S.n(), S.kernels.data(), S.biases.data(), S.bn_scale.data(),
                                           S.bn_shift.data(), S.ksize.data(), S.cin.data(), S.cout.data(), S.flags.data(), b.act, b.pad, b.pad_value});
            small_out[&b] = y;
        }
        if (!descs.empty()) RUN(c, sms::smallmap_stack_multi(descs.data(), (int)descs.size(), B, c.st));
    }
    for (const BlockCfg& b : h.blocks) {
        const std::string name = "hpnn/" + b.name;
        Branch r{&b, false, B8(), F32()};
        if (branch_on_tc(b)) {
            B8 hb;
            TRY(to_blk8(c, pools[b.ds], Bcap, bsplit, b.pad, nullptr, 0, &hb));
            TcArgs a; a.act = b.act; a.pad = b.pad; a.next_pad = b.pad;
            B8 y;
            TRY(conv_tc(c, hb, name + "/conv0", a, Bcap, &y));
            c.free(hb);
            hb = y;
            for (int rr = 1; rr < b.n_convs; ++rr) {
                TRY(resnet_tc(c, hb, name + "/resnet" + std::to_string(rr), b.act, b.pad, b.use_bn, nullptr,
                              rr + 1 < b.n_convs ? b.pad : PCNN_PAD_CONSTANT, Bcap, &y));
                hb = y;
            }
            if (um_tc && b.deconv) { r.is_b8 = true; r.b8 = hb; }
            else { TRY(from_blk8(c, hb, hb.C, Bcap, nullptr, &r.f32)); c.free(hb); }      // resize branches read NCHW fp32
        } else {
            auto done = small_out.find(&b);
            if (done != small_out.end()) r.f32 = done->second;
            else TRY(branch_lowres_f32(c, b, pools[b.ds], Bcap, &r.f32));
            if (um_tc && b.deconv) {      // tiny map (< 16 pixels a side) computed by the FP32 stack kernel
                TRY(to_blk8(c, r.f32, Bcap, 1, PCNN_PAD_CONSTANT, nullptr, 0, &r.b8));
                c.free(r.f32);
                r.is_b8 = true;
            }
        }
        br.push_back(r);
    }
    for (auto& kv : pools) c.free(kv.second);

    B8 cat = c.b8(Bcap, B, 2 * F, H, W, split, false);
    {
        TcArgs a; a.act = PCNN_ACT_LEAKY_RELU; a.into = &cat;
        B8 dummy;
        TRY(conv_tc(c, x0, "hpnn/non_bottleneck_conv", a, Bcap, &dummy));
        c.free(x0);
    }
    const bool fused_dc = F % 8 == 0 && ndeconv <= 8 && (int)h.blocks.size() - ndeconv <= 8 && strides_ok;
    bool fused = fused_dc;
    for (const Branch& r : br)
        if (!r.cfg->deconv) fused = fused && (size_t)F * r.f32.H * r.f32.W <= 8192;
    const bool two_pass = um_tc == 2 && fused_dc;
    if (um_tc == 1 && !fused) um_tc = 0;          // (sources between the two limits: general kernels, as before)
    if (um_tc == 2 && !fused_dc) um_tc = 0;
    if (two_pass) fused = true;
    {
        // operand lists of the fused kernels (host arrays of device pointers)
        std::vector<const void*> d_in, d_k;
        std::vector<const float*> d_b, r_in, r_wy, r_wx;
        std::vector<const int32_t*> r_iy, r_ix;
        std::vector<int> d_s, d_ih, d_iw, d_act, r_t, r_ih, r_iw;
        if (fused) {
            for (Branch& r : br) {
                const std::string n = "hpnn/" + r.cfg->name + "/deconv";
                if (r.cfg->deconv) {
                    if (!um_tc && r.is_b8) { TRY(from_blk8(c, r.b8, r.b8.C, Bcap, nullptr, &r.f32)); c.free(r.b8); r.is_b8 = false; }
                    const Weight* bias = m.find(n + "/bias");
                    d_in.push_back(um_tc ? (const void*)r.b8.hi : (const void*)r.f32.p);
                    d_k.push_back(um_tc ? m.deconv_tc[n] : m.deconv_f32[n]);
                    d_b.push_back(bias ? bias->dev : nullptr);
                    d_s.push_back(r.cfg->us);
                    d_ih.push_back(um_tc ? r.b8.H : r.f32.H);
                    d_iw.push_back(um_tc ? r.b8.W : r.f32.W);
                    d_act.push_back(r.cfg->deconv_act);
                } else {
                    DevAxis ay, ax;
                    TRY(axis_table(c, r.f32.H, H, r.cfg->resize_method, &ay));
                    TRY(axis_table(c, r.f32.W, W, r.cfg->resize_method, &ax));
                    r_in.push_back(r.f32.p); r_iy.push_back(ay.idx); r_wy.push_back(ay.w); r_ix.push_back(ax.idx); r_wx.push_back(ax.w);
                    r_t.push_back(ay.taps); r_ih.push_back(r.f32.H); r_iw.push_back(r.f32.W);
                }
            }
            const void* nullp = nullptr; const float* nullf = nullptr; const int32_t* nulli = nullptr; int zero = 0;
            auto P = [&](auto& v, auto& dflt) { return v.empty() ? &dflt : v.data(); };
            if (um_tc) {
                const int nrs = two_pass ? 0 : (int)r_in.size();
                RUN(c, pcnn_upsample_merge_tc_blk8((int)d_in.size(), P(d_in, nullp), P(d_k, nullp), P(d_b, nullf), P(d_s, zero), P(d_ih, zero), P(d_iw, zero),
                                                   P(d_act, zero), nrs, P(r_in, nullf), P(r_iy, nulli), P(r_wy, nullf), P(r_ix, nulli),
                                                   P(r_wx, nullf), P(r_t, zero), P(r_ih, zero), P(r_iw, zero), alpha, cat.hi, cat.lo, cat.mode, B, H, W, cat.C, F, c.st));
                if (two_pass)
                    RUN(c, pcnn_resize_add_blk8((int)r_in.size(), r_in.data(), r_iy.data(), r_wy.data(), r_ix.data(), r_wx.data(), r_t.data(),
                                                r_ih.data(), r_iw.data(), alpha, cat.hi, cat.lo, cat.mode, B, F, H, W, cat.C, F, c.st));
            } else
                RUN(c, pcnn_upsample_merge_blk8((int)d_in.size(), reinterpret_cast<const float* const*>(P(d_in, nullp)),
                                                reinterpret_cast<const float* const*>(P(d_k, nullp)), P(d_b, nullf), P(d_s, zero), P(d_ih, zero),
                                                P(d_iw, zero), P(d_act, zero), (int)r_in.size(), P(r_in, nullf), P(r_iy, nulli), P(r_wy, nullf),
                                                P(r_ix, nulli), P(r_wx, nullf), P(r_t, zero), P(r_ih, zero), P(r_iw, zero), alpha, cat.hi, cat.lo,
                                                cat.mode, B, F, H, W, cat.C, F, c.st));
        } else {        // general kernels: fp32 merge buffer, one read-modify-write per branch
            F32 merged = c.f32(Bcap, B, F, H, W);
            bool first = true;
            for (Branch& r : br) {
                if (!r.cfg->deconv) continue;
                if (r.is_b8) { TRY(from_blk8(c, r.b8, r.b8.C, Bcap, nullptr, &r.f32)); c.free(r.b8); r.is_b8 = false; }
                const std::string n = "hpnn/" + r.cfg->name + "/deconv";
                const Weight* k = m.find(n + "/kernel");
                const Weight* bias = m.find(n + "/bias");
                RUN(c, pcnn_deconv_same_f32(r.f32.p, k->dev, bias ? bias->dev : nullptr, merged.p, B, F, F, r.f32.H, r.f32.W, H, W, r.cfg->us,
                                            r.cfg->us, r.cfg->us, r.cfg->deconv_act, alpha, first ? 0 : 1, merged.bs, c.st));
                first = false;
            }
            for (Branch& r : br) {
                if (r.cfg->deconv) continue;
                DevAxis ay, ax;
                TRY(axis_table(c, r.f32.H, H, r.cfg->resize_method, &ay));
                TRY(axis_table(c, r.f32.W, W, r.cfg->resize_method, &ax));
                RUN(c, pcnn_resize_f32(r.f32.p, ay.idx, ay.w, ax.idx, ax.w, ay.taps, merged.p, B, F, r.f32.H, r.f32.W, H, W, alpha, first ? 0 : 1, merged.bs, c.st));
                first = false;
            }
            TRY(to_blk8(c, merged, Bcap, cat.mode, PCNN_PAD_CONSTANT, &cat, F, nullptr));
            c.free(merged);
        }
        if (cat.halo.mode != PCNN_PAD_CONSTANT) cat.halo.pad = -1;
    }
    for (Branch& r : br) { if (r.is_b8) c.free(r.b8); else c.free(r.f32); }

    B8 y;
    {
        TcArgs a; a.act = PCNN_ACT_LEAKY_RELU;
        TRY(conv_tc(c, cat, "hpnn/post_merge_conv", a, Bcap, &y));
        c.free(cat);
    }
    float* d;
    TRY(dx_mlp(c, dx, B, Bcap, H, W, &d));
    {
        B8 nxt;
        TRY(resnet_tc(c, y, "hpnn/post_merge_resnet", PCNN_ACT_LEAKY_RELU, PCNN_PAD_CONSTANT, false, d, PCNN_PAD_CONSTANT, Bcap, &nxt));
        y = nxt;
    }
    const int S = (int)h.fin.filters.size(), nreg = h.fin.nreg;
    for (int k = 0; k < S - nreg; ++k) {
        const std::string n = "hpnn/final/" + std::to_string(k);
        TcArgs a; a.act = h.fin.act; a.pad = h.fin.pad;
        B8 t2, nxt;
        TRY(conv_tc(c, y, n + "/conv", a, Bcap, &t2));
        c.free(y);
        TRY(resnet_tc(c, t2, n + "/resnet", h.fin.act, PCNN_PAD_CONSTANT, false, nullptr, PCNN_PAD_CONSTANT, Bcap, &nxt));
        y = nxt;
    }
    for (int k = S - nreg; k < S; ++k) {      // the last linear convs: 16 output rows x 8 channel slots per tile
        TcArgs a; a.act = PCNN_ACT_LINEAR;
        B8 nxt;
        TRY(conv_tc(c, y, "hpnn/final/" + std::to_string(k) + "/conv", a, Bcap, &nxt));
        c.free(y);
        y = nxt;
    }
    F32 cat2, yf;
    bool in_cat2 = false;
    if (h.use_scaling && y.C == 1) {
        cat2 = c.f32(Bcap, B, 2, H, W);
        F32 view = cat2; view.slot = -1; view.C = 1;
        TRY(from_blk8(c, y, 1, Bcap, &view, &yf));
        in_cat2 = true;
    } else {
        TRY(from_blk8(c, y, y.C, Bcap, nullptr, &yf));
    }
    c.free(y);
    return hpnn_tail(c, yf, in_cat2, cat2, rhs, dx, S, B, Bcap, H, W, out);
}

static int hpnn_run(Ctx& c, const float* rhs, const float* dx, float* out, int B, int Bcap, int H, int W) {
    return c.m->hp_mode == 0 ? hpnn_fp32(c, rhs, dx, out, B, Bcap, H, W) : hpnn_tc(c, rhs, dx, out, B, Bcap, H, W);
}

// ------------------------------------------------------------------------------------------------ DBCNN
// The separable first 2-D convolution's per-row weights (see pcnn.h, pcnn_conv2d_tc_rowweights), built on the host:
// A_x[b,m,co] = sum_a W[a,b,m,co] * S[m, x+a-k/2], packed as [ceil(Cin/16)][k][2][T][cp][8] fp16 with a power-of-two pre-scale.
struct RowWeights { std::vector<uint16_t> img; float acc_scale = 1.f; };
static RowWeights build_rowweights(const Weight& kern, int M, int xres, int cp, int rt, int T) {
    const int k = (int)kern.shape[0], Cin = (int)kern.shape[2], Cout = (int)kern.shape[3], Hh = xres, pad = k / 2;
    std::vector<float> basis((size_t)Cin * Hh);
    {
        const std::vector<float> sb = sinh_basis(M, xres), px = position_table(xres);
        std::copy(sb.begin(), sb.end(), basis.begin());
        for (int x = 0; x < Hh; ++x) { basis[(size_t)M * Hh + x] = px[x]; basis[(size_t)(M + 1) * Hh + x] = 1.0f; }
    }
    std::vector<float> A((size_t)Hh * k * Cin * Cout);
    float amax = 0.f;
    for (int x = 0; x < Hh; ++x)
        for (int b = 0; b < k; ++b)
            for (int mch = 0; mch < Cin; ++mch)
                for (int co = 0; co < Cout; ++co) {
                    double acc = 0.0;
                    for (int a = 0; a < k; ++a) {
                        const int xs = x + a - pad;
                        if (xs < 0 || xs >= Hh) continue;
                        acc += (double)basis[(size_t)mch * Hh + xs] * (double)kern.host[(((size_t)a * k + b) * Cin + mch) * Cout + co];
                    }
                    const float v = (float)acc;
                    A[(((size_t)x * k + b) * Cin + mch) * Cout + co] = v;
                    amax = std::max(amax, std::fabs(v));
                }
    const float scale = pow2_prescale(amax);
    const int c16 = cdiv(Cin, 16);
    RowWeights rw;
    rw.acc_scale = 1.0f / scale;
    rw.img.assign((size_t)c16 * k * 2 * T * cp * 8, 0);
    for (int cc = 0; cc < c16; ++cc)
        for (int b = 0; b < k; ++b)
            for (int half = 0; half < 2; ++half)
                for (int t = 0; t < T; ++t) {
                    const int x = (t / rt) * rt + (rt - 1 - t % rt);
                    if (x >= Hh) continue;
                    for (int co = 0; co < Cout && co < cp; ++co)
                        for (int j = 0; j < 8; ++j) {
                            const int mch = cc * 16 + half * 8 + j;
                            if (mch >= Cin) continue;
                            const __half hv = __float2half_rn(A[(((size_t)x * k + b) * Cin + mch) * Cout + co] * scale);
                            uint16_t bits;
                            std::memcpy(&bits, &hv, 2);
                            rw.img[(((((size_t)cc * k + b) * 2 + half) * T + t) * cp + co) * 8 + j] = bits;
                        }
                }
    return rw;
}

// everything up to (not including) the final max-normalisation: raw [B,1,xres,n]
static int dbcnn_raw(Ctx& c, const float* bc, const float* dx, int B, int Bcap, int n, int xres, F32* raw_out) {
    Model& m = *c.m;
    const DbcnnCfg& d = m.db;
    const float* posy;
    TRY(pos_table(c, n, &posy));
    F32 hcur = c.f32(Bcap, B, 3, 1, n);
    RUN(c, pcnn_dbcnn_input_f32(bc, 1.0f /* cos(pi * linspace(0,1,xres)[0]) */, posy, hcur.p, B, n, c.st));
    {
        std::vector<std::string> dummy;
        StackLayers L;
        const int nb = (int)d.bnd.filters.size();
        for (int k = 0; k < nb; ++k) {
            const std::string base = "dbcnn/boundary/" + std::to_string(k), bn = base + "/bn";
            StackLayers one;
            TRY(stack_layers(m, base + "/conv", {base + "/resnet"}, 1, d.use_bn ? bn.c_str() : nullptr, d.use_bn, &one));
            L.kernels.insert(L.kernels.end(), one.kernels.begin(), one.kernels.end());
            L.biases.insert(L.biases.end(), one.biases.begin(), one.biases.end());
            L.bn_scale.insert(L.bn_scale.end(), one.bn_scale.begin(), one.bn_scale.end());
            L.bn_shift.insert(L.bn_shift.end(), one.bn_shift.begin(), one.bn_shift.end());
            L.ksize.insert(L.ksize.end(), one.ksize.begin(), one.ksize.end());
            L.cin.insert(L.cin.end(), one.cin.begin(), one.cin.end());
            L.cout.insert(L.cout.end(), one.cout.begin(), one.cout.end());
            L.flags.insert(L.flags.end(), one.flags.begin(), one.flags.end());
        }
        if (boundary_supported(n, L)) {
            // the whole 1-D stack (n_boundary x (conv [+BN] + resnet)) in ONE kernel, activations in shared memory
            F32 y = c.f32(Bcap, B, L.cout.back(), 1, n);
            RUN(c, pcnn_boundary_stack_f32(hcur.p, y.p, B, n, 3, L.n(), L.kernels.data(), L.biases.data(), L.bn_scale.data(), L.bn_shift.data(),
                                           L.ksize.data(), L.cin.data(), L.cout.data(), L.flags.data(), d.bnd.act, d.bnd.pad, d.bnd.pad_value, c.st));
            c.free(hcur);
            hcur = y;
        } else {
            for (int k = 0; k < nb; ++k) {
                const std::string base = "dbcnn/boundary/" + std::to_string(k);
                ConvW w;
                TRY(m.conv(base + "/conv", 1, &w));
                if (d.use_bn) m.with_bn(base + "/bn", &w);
                F32 t, nxt;
                TRY(conv_f32(c, hcur, w, d.bnd.act, d.bnd.pad, d.bnd.pad_value, nullptr, nullptr, &t, Bcap));
                c.free(hcur);
                TRY(resnet_f32(c, t, base + "/resnet", 1, d.bnd.act, d.bnd.pad, d.bnd.pad_value, d.use_bn, nullptr, Bcap, &nxt));
                hcur = nxt;
            }
        }
    }
    const int M = hcur.C;
    // SPP over the boundary features + domain info -> MLP -> mode weights
    const int nbins = spp_bins(d.spp_levels, 1);
    const int32_t* boxes;
    const std::vector<std::vector<int>>& lv = d.spp_levels;
    TRY((table<int32_t>(c, "sppbox1/" + std::to_string(n), (size_t)nbins * 4, [&lv, n] { return spp_boxes(lv, 1, n, 1); }, &boxes)));
    float* spp = c.vec((size_t)Bcap * nbins);
    RUN(c, pcnn_spp_f32(hcur.p, boxes, spp, B, M, 1, n, nbins, d.spp_mode, c.st));
    float* v = c.vec((size_t)Bcap * (3 + nbins));
    RUN(c, pcnn_dense_input_f32(dx, spp, v, B, xres, n, nbins, 1, c.st));
    for (size_t i = 0; i < d.mlp_units.size(); ++i) {
        float* nxt;
        int nout;
        TRY(dense(c, v, B, Bcap, "dbcnn/mlp/" + std::to_string(i), d.mlp_acts[i], &nxt, &nout));
        v = nxt;
    }
    const int S = (int)d.fin.filters.size(), nreg = d.fin.nreg;
    const float *posx = nullptr, *sbasis = nullptr;
    TRY(pos_table(c, xres, &posx));
    TRY((table<float>(c, "sinh/" + std::to_string(M) + "/" + std::to_string(xres), (size_t)M * xres, [M, xres] { return sinh_basis(M, xres); }, &sbasis)));
    if (m.db_mode != 0) {
        const bool sep = m.db_mode == 1 && S - nreg >= 1;
        B8 t;
        const uint16_t* rw_img = nullptr;
        float rw_scale = 1.f;
        int rwk = 0, rwcin = 0, rwcout = 0;
        if (sep) {
            // the first 2-D convolution sees a SEPARABLE input: its row taps fold into per-row weights, the [B,29,xres,n]
            // expansion is never written (14 MMAs per tile, not 121)
            const Weight* kern = m.find("dbcnn/final/0/conv/kernel");
            rwk = (int)kern->shape[0]; rwcin = (int)kern->shape[2]; rwcout = (int)kern->shape[3];
            const int cp = pcnn_conv_tc_channel_slots(rwcout, rwk), T = pcnn_conv_tc_rowweight_slots(rwcout, rwk, xres);
            if (cp == 0 || T == 0) { set_error("pack_rowweights_tc: unsupported layer (Cout <= 32, odd k <= 15)"); return PCNN_ERR_UNSUPPORTED; }
            const int rt = cp == 24 ? 5 : 128 / cp;
            const size_t count = (size_t)cdiv(rwcin, 16) * rwk * 2 * T * cp * 8;
            // the power-of-two pre-scale comes with the image: it is kept in the Prepared record next to the table
            const std::string key = "rowweights/" + std::to_string(xres);
            Prepared* prep = c.prep;
            TRY((table<uint16_t>(c, key, count, [&] { RowWeights r = build_rowweights(*kern, M, xres, cp, rt, T); prep->scalars[key] = r.acc_scale; return r.img; }, &rw_img)));
            if (!c.dry) rw_scale = prep->scalars[key];
            t = c.b8(Bcap, B, M + 2, 1, n, 1, false);
            RUN(c, pcnn_dbcnn_signal_blk8(hcur.p, v, posy, t.hi, B, M, n, c.st));
            if (t.halo.mode != PCNN_PAD_CONSTANT) t.halo.pad = -1;
        } else {
            t = c.b8(Bcap, B, M + 2, xres, n, m.db_mode, false);
            RUN(c, pcnn_dbcnn_expand_blk8(hcur.p, sbasis, v, posx, posy, t.hi, t.lo, t.mode, B, M, xres, n, c.st));
            if (t.halo.mode != PCNN_PAD_CONSTANT) t.halo.pad = -1;
        }
        c.free(hcur);
        for (int k = 0; k < S - nreg; ++k) {
            const std::string nme = "dbcnn/final/" + std::to_string(k);
            B8 y;
            if (sep && k == 0) {
                TRY(halo_fill(c, t, rwk / 2, PCNN_PAD_CONSTANT));
                y = c.b8(Bcap, B, rwcout, xres, n, 1, false);
                const Weight* bias = m.find(nme + "/conv/bias");
                RUN(c, pcnn_conv2d_tc_rowweights(t.hi, rw_img, bias ? bias->dev : nullptr, y.hi, B, rwcin, rwcout, y.C, xres, n, rwk, d.fin.act,
                                                 rw_scale, c.num_sms, c.st));
                if (y.halo.mode != PCNN_PAD_CONSTANT) y.halo.pad = -1;
            } else {
                TcArgs a; a.act = d.fin.act;
                TRY(conv_tc(c, t, nme + "/conv", a, Bcap, &y));
            }
            c.free(t);
            TRY(resnet_tc(c, y, nme + "/resnet", d.fin.act, PCNN_PAD_CONSTANT, false, nullptr, PCNN_PAD_CONSTANT, Bcap, &t));
        }
        for (int k = S - nreg; k < S; ++k) {
            TcArgs a; a.act = PCNN_ACT_TANH;
            B8 y;
            TRY(conv_tc(c, t, "dbcnn/final/" + std::to_string(k) + "/conv", a, Bcap, &y));
            c.free(t);
            t = y;
        }
        TRY(from_blk8(c, t, t.C, Bcap, nullptr, raw_out));
        c.free(t);
        return PCNN_OK;
    }
    F32 o = c.f32(Bcap, B, M + 2, xres, n);
    RUN(c, pcnn_dbcnn_expand_f32(hcur.p, sbasis, v, posx, posy, o.p, B, M, xres, n, c.st));
    c.free(hcur);
    for (int k = 0; k < S - nreg; ++k) {
        const std::string nme = "dbcnn/final/" + std::to_string(k);
        ConvW w;
        TRY(m.conv(nme + "/conv", 2, &w));
        F32 t, nxt;
        TRY(conv_f32(c, o, w, d.fin.act, d.fin.pad, d.fin.pad_value, nullptr, nullptr, &t, Bcap));
        c.free(o);
        TRY(resnet_f32(c, t, nme + "/resnet", 2, d.fin.act, PCNN_PAD_CONSTANT, 0.f, false, nullptr, Bcap, &nxt));
        o = nxt;
    }
    for (int k = S - nreg; k < S; ++k) {
        ConvW w;
        TRY(m.conv("dbcnn/final/" + std::to_string(k) + "/conv", 2, &w));
        F32 t;
        TRY(conv_f32(c, o, w, PCNN_ACT_TANH, PCNN_PAD_CONSTANT, 0.f, nullptr, nullptr, &t, Bcap));
        c.free(o);
        o = t;
    }
    *raw_out = o;
    return PCNN_OK;
}

static int dbcnn_run(Ctx& c, const float* bc, const float* dx, float* out, int B, int Bcap, int n, int xres) {
    const DbcnnCfg& d = c.m->db;
    F32 raw;
    TRY(dbcnn_raw(c, bc, dx, B, Bcap, n, xres, &raw));
    if (raw.C != 1) { set_error("the DBCNN's last convolution must produce one channel, got %d", raw.C); return PCNN_ERR_INVALID_ARGUMENT; }
    float* mx = c.vec((size_t)Bcap);
    RUN(c, pcnn_maxabs_f32(raw.p, mx, B, (int64_t)xres * n, c.st));
    if (d.postsmooth > 0) {
        float* pre = c.vec((size_t)Bcap * xres * n);
        float* zero = c.vec((size_t)Bcap * xres * n);
        if (!c.dry) PCNN_CHECK_CUDA(cudaMemsetAsync(zero, 0, (size_t)B * xres * n * 4, c.st));
        RUN(c, pcnn_dbcnn_finalize_f32(raw.p, mx, bc, pre, B, xres, n, c.st));
        TRY(jacobi(c, pre, zero, dx, B, Bcap, xres, n, d.postsmooth, out));
    } else {
        RUN(c, pcnn_dbcnn_finalize_f32(raw.p, mx, bc, out, B, xres, n, c.st));
    }
    c.free(raw);
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------------ Poisson_CNN_Legacy
// models/Poisson_CNN_Legacy.py:15-51: normalise 5 inputs, HPNN, 4 x DBCNN (batched: the weights are shared), oriented merge
static int pcnn_run(Ctx& c, const float* rhs, const float* left, const float* top, const float* right, const float* bottom,
                    const float* dx, float* out, int B, int Bcap, int nx, int ny, int jacobi_iters) {
    float* mrhs = c.vec(Bcap); float* ml = c.vec(Bcap); float* mt = c.vec(Bcap); float* mr = c.vec(Bcap); float* mb = c.vec(Bcap);
    float* rhs_n = c.vec((size_t)Bcap * nx * ny);
    RUN(c, pcnn_maxabs_f32(rhs, mrhs, B, (int64_t)nx * ny, c.st));
    RUN(c, pcnn_scale_inv_f32(rhs, mrhs, rhs_n, B, (int64_t)nx * ny, c.st));
    RUN(c, pcnn_maxabs_f32(left, ml, B, ny, c.st));
    RUN(c, pcnn_maxabs_f32(top, mt, B, nx, c.st));
    RUN(c, pcnn_maxabs_f32(right, mr, B, ny, c.st));
    RUN(c, pcnn_maxabs_f32(bottom, mb, B, nx, c.st));
    float* hp = c.vec((size_t)Bcap * nx * ny);
    TRY(hpnn_run(c, rhs_n, dx, hp, B, Bcap, nx, ny));
    const size_t plane = (size_t)nx * ny;
    const float *L, *T, *R, *Bt;
    // The DBCNN's weights are shared by the four boundaries, so their problems can be batched: all four in one call for
    // small batches (fills the GPU at batch 1), two per call in between, one per call once a single boundary already is
    // >= 64 problems of 256x256 (thousands of tiles per launch).  Larger calls save the fixed cost of ~25 launches each, but
    // the DBCNN's activations then are 4x / 2x the HPNN's.  Measured at the headline shape (B = 256 in 128-sample slices,
    // back-to-back runs on one box, power-capped): 4B 1176, 2B 1142, B 1144 solutions/s -- inside the run-to-run spread, so
    // the smaller workspace wins (12.2 GB instead of 17.6 / ~20 GB).
    static const long long side_px = std::getenv("PCNN_ENGINE_BATCH_SIDES_PX") ? std::atoll(std::getenv("PCNN_ENGINE_BATCH_SIDES_PX")) : 64LL * 65536;
    const int group = ((long long)4 * Bcap * plane <= side_px) ? 4 : (((long long)2 * Bcap * plane <= side_px) ? 2 : 1);
    float* dxr = c.vec((size_t)4 * Bcap);
    for (int i = 0; i < 4; ++i)
        if (!c.dry) PCNN_CHECK_CUDA(cudaMemcpyAsync(dxr + (size_t)i * B, dx, (size_t)B * 4, cudaMemcpyDeviceToDevice, c.st));
    if (nx == ny) {
        float* bcs = c.vec((size_t)4 * Bcap * ny);
        float* res = c.vec((size_t)4 * Bcap * plane);
        const float* src[4] = {left, top, right, bottom};
        const float* mm[4] = {ml, mt, mr, mb};
        for (int i = 0; i < 4; ++i) RUN(c, pcnn_scale_inv_f32(src[i], mm[i], bcs + (size_t)i * B * ny, B, ny, c.st));
        for (int i = 0; i < 4; i += group)
            TRY(dbcnn_run(c, bcs + (size_t)i * B * ny, dxr, res + (size_t)i * B * plane, group * B, group * Bcap, ny, nx));
        L = res; T = res + (size_t)B * plane; R = res + (size_t)2 * B * plane; Bt = res + (size_t)3 * B * plane;
    } else {
        float* lr = c.vec((size_t)2 * Bcap * ny);
        float* tb = c.vec((size_t)2 * Bcap * nx);
        float* res_lr = c.vec((size_t)2 * Bcap * plane);
        float* res_tb = c.vec((size_t)2 * Bcap * plane);
        RUN(c, pcnn_scale_inv_f32(left, ml, lr, B, ny, c.st));
        RUN(c, pcnn_scale_inv_f32(right, mr, lr + (size_t)B * ny, B, ny, c.st));
        RUN(c, pcnn_scale_inv_f32(top, mt, tb, B, nx, c.st));
        RUN(c, pcnn_scale_inv_f32(bottom, mb, tb + (size_t)B * nx, B, nx, c.st));
        const int g2 = group >= 2 ? 2 : 1;
        for (int i = 0; i < 2; i += g2) TRY(dbcnn_run(c, lr + (size_t)i * B * ny, dxr, res_lr + (size_t)i * B * plane, g2 * B, g2 * Bcap, ny, nx));
        for (int i = 0; i < 2; i += g2) TRY(dbcnn_run(c, tb + (size_t)i * B * nx, dxr, res_tb + (size_t)i * B * plane, g2 * B, g2 * Bcap, nx, ny));
        L = res_lr; R = res_lr + (size_t)B * plane; T = res_tb; Bt = res_tb + (size_t)B * plane;
    }
    if (jacobi_iters > 0) {
        // the reference passes the max-normalised rhs here (it rebinds `rhs`, Poisson_CNN_Legacy.py:23,49)
        float* pre = c.vec((size_t)Bcap * plane);
        RUN(c, pcnn_merge_f32(hp, L, T, R, Bt, dx, mrhs, ml, mt, mr, mb, pre, B, nx, ny, c.st));
        TRY(jacobi(c, pre, rhs_n, dx, B, Bcap, nx, ny, jacobi_iters, out));
    } else {
        RUN(c, pcnn_merge_f32(hp, L, T, R, Bt, dx, mrhs, ml, mt, mr, mb, out, B, nx, ny, c.st));
    }
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------------ sliced execution
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; ok = false; cudaGetLastError(); return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) { ok = false; cudaGetLastError(); }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int slice_capacity(const Model& m, int B, long long pixels, int unit) {
    long long mb = m.microbatch > 0 ? m.microbatch : std::max<long long>(1, 128LL * 65536 / std::max<long long>(pixels, 1));
    mb = std::min<long long>(mb, 65535 / std::max(unit, 1));       // several kernels put the batch into gridDim.y / z
    return (int)std::min<long long>(mb, B);
}

// body(ctx, first sample, samples in this slice, slice capacity)
template <typename Body>
static int run_sliced(Model& m, const std::string& kind, int B, int d0, int d1, int unit, void* ws, size_t ws_bytes,
                      cudaStream_t st, bool query_only, size_t* need_out, Body body) {
    if (!m.finalized) { set_error("call pcnn_finalize_weights() first"); return PCNN_ERR_INVALID_ARGUMENT; }
    if (B <= 0 || d0 <= 0 || d1 <= 0) { set_error("batch and grid sizes must be positive"); return PCNN_ERR_INVALID_ARGUMENT; }
    const int cap = slice_capacity(m, B, (long long)d0 * d1, unit);
    const std::string key = kind + ":" + std::to_string(cap) + ":" + std::to_string(d0) + ":" + std::to_string(d1);
    size_t need;
    auto it = m.need_cache.find(key);
    if (it == m.need_cache.end()) {
        Prepared tmp;
        Ctx c;
        c.m = &m; c.dry = true; c.prep = &tmp; c.num_sms = m.num_sms;
        TRY(body(c, 0, cap, cap));
        need = tmp.top;
        m.need_cache[key] = need;
    } else {
        need = it->second;
    }
    if (need_out) *need_out = need;
    if (query_only) return PCNN_OK;
    if (!ws) { set_error("workspace is null (%zu bytes needed: pcnn_workspace_bytes)", need); return PCNN_ERR_INVALID_ARGUMENT; }
    if (ws_bytes < need) { set_error("workspace holds %zu bytes, this call needs %zu (pcnn_workspace_bytes)", ws_bytes, need); return PCNN_ERR_INVALID_ARGUMENT; }
    if (m.prepared.size() > 16 && !m.prepared.count(ws)) m.prepared.clear();
    Prepared& P = m.prepared[ws];
    if (!P.valid || P.shape_key != key) {
        // first use of this workspace for this shape: BLK8 halos and channel padding must start as zeros
        P = Prepared();
        P.shape_key = key;
        P.valid = true;
        PCNN_CHECK_CUDA(cudaMemsetAsync(ws, 0, need, st));
    }
    Ctx c;
    c.m = &m; c.dry = false; c.base = reinterpret_cast<char*>(ws); c.st = st; c.prep = &P; c.num_sms = m.num_sms;
    for (int lo = 0; lo < B; lo += cap) {
        P.reset();
        const int st_ = body(c, lo, std::min(cap, B - lo), cap);
        if (st_ != PCNN_OK) { P.valid = false; return st_; }
        if (P.top > need) { P.valid = false; set_error("internal: arena grew to %zu bytes beyond the planned %zu", P.top, need); return PCNN_ERR_CUDA; }
    }
    return PCNN_OK;
}

static int hpnn_entry(Model& m, const float* rhs, const float* dx, float* out, int B, int H, int W, void* ws, size_t ws_bytes,
                      cudaStream_t st, bool query, size_t* need) {
    if (!m.hp.present) { set_error("this handle has no hpnn_model"); return PCNN_ERR_INVALID_ARGUMENT; }
    return run_sliced(m, "hpnn", B, H, W, 1, ws, ws_bytes, st, query, need, [&](Ctx& c, int lo, int nb, int cap) {
        const size_t plane = (size_t)H * W;
        return hpnn_run(c, rhs ? rhs + lo * plane : nullptr, dx ? dx + lo : nullptr, out ? out + lo * plane : nullptr, nb, cap, H, W);
    });
}

static int dbcnn_entry(Model& m, const float* bc, const float* dx, float* out, int B, int n, int xres, void* ws, size_t ws_bytes,
                       cudaStream_t st, bool query, size_t* need) {
    if (!m.db.present) { set_error("this handle has no dbcnn_model"); return PCNN_ERR_INVALID_ARGUMENT; }
    return run_sliced(m, "dbcnn", B, xres, n, 1, ws, ws_bytes, st, query, need, [&](Ctx& c, int lo, int nb, int cap) {
        return dbcnn_run(c, bc ? bc + (size_t)lo * n : nullptr, dx ? dx + lo : nullptr, out ? out + (size_t)lo * xres * n : nullptr, nb, cap, n, xres);
    });
}

static int pcnn_entry(Model& m, int jacobi_iters, const float* rhs, const float* left, const float* top, const float* right,
                      const float* bottom, const float* dx, float* out, int B, int nx, int ny, void* ws, size_t ws_bytes,
                      cudaStream_t st, bool query, size_t* need) {
    if (!m.hp.present || !m.db.present) { set_error("pcnn_forward needs a handle with both hpnn_model and dbcnn_model"); return PCNN_ERR_INVALID_ARGUMENT; }
    return run_sliced(m, "pcnn", B, nx, ny, 4, ws, ws_bytes, st, query, need, [&](Ctx& c, int lo, int nb, int cap) {
        const size_t plane = (size_t)nx * ny;
        auto off = [&](const float* p, size_t stride) { return p ? p + lo * stride : nullptr; };
        return pcnn_run(c, off(rhs, plane), off(left, ny), off(top, nx), off(right, ny), off(bottom, nx), off(dx, 1),
                        out ? out + lo * plane : nullptr, nb, cap, nx, ny, jacobi_iters);
    });
}

struct Handle {
    Model model;
    int jacobi_iterations = 0;
};

}  // namespace eng
}  // namespace pcnn

using namespace pcnn;
using pcnn::eng::Handle;

struct pcnn_model { Handle h; };

#define ENG_GUARD_BEGIN try {
#define ENG_GUARD_END                                                            \
    } catch (const std::exception& e) {                                          \
        ::pcnn::set_error("%s", e.what());                                       \
        return PCNN_ERR_INVALID_ARGUMENT;                                        \
    } catch (...) {                                                              \
        ::pcnn::set_error("unknown C++ exception");                              \
        return PCNN_ERR_CUDA;                                                    \
    }

extern "C" int pcnn_create(const char* config_json, int device, pcnn_handle* out) {
    PCNN_CHECK_ARG(config_json && out, "pcnn_create: null argument");
    ENG_GUARD_BEGIN
    const json::Value cfg = json::parse(config_json);
    if (cfg.type != json::Value::Obj) { set_error("pcnn_create: the config must be a JSON object"); return PCNN_ERR_INVALID_ARGUMENT; }
    std::unique_ptr<pcnn_model> pm(new pcnn_model());
    Handle& h = pm->h;
    h.model.device = device;
    const json::Value* hp = cfg.find("hpnn_model");
    if (!hp) hp = cfg.find("model");
    const json::Value* db = cfg.find("dbcnn_model");
    if (!hp && !db) { set_error("pcnn_create: the config needs an 'hpnn_model' (or 'model') and/or a 'dbcnn_model' section"); return PCNN_ERR_INVALID_ARGUMENT; }
    if (hp) h.model.hp = eng::parse_hpnn(*hp);
    if (db) h.model.db = eng::parse_dbcnn(*db);
    h.jacobi_iterations = (int)cfg.number("jacobi_iterations", 0);
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) h.model.num_sms = sms;
    else cudaGetLastError();
    *out = pm.release();
    return PCNN_OK;
    ENG_GUARD_END
}

extern "C" int pcnn_destroy(pcnn_handle handle) {
    if (!handle) return PCNN_OK;
    eng::DeviceGuard g(handle->h.model.device);
    delete handle;
    return PCNN_OK;
}

extern "C" int pcnn_set_weight(pcnn_handle handle, const char* name, const void* host_ptr, const int64_t* shape, int ndim, int dtype) {
    PCNN_CHECK_ARG(handle && name && host_ptr && shape && ndim >= 1 && ndim <= 4, "pcnn_set_weight: bad argument");
    PCNN_CHECK_ARG(dtype == 0 || dtype == 1, "pcnn_set_weight: dtype must be 0 (float32) or 1 (float64)");
    ENG_GUARD_BEGIN
    eng::Model& m = handle->h.model;
    eng::DeviceGuard g(m.device);
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) { PCNN_CHECK_ARG(shape[i] > 0, "pcnn_set_weight: '%s' has a non-positive dimension", name); n *= (size_t)shape[i]; }
    eng::Weight& w = m.w[name];
    w.shape.assign(shape, shape + ndim);
    w.host.resize(n);
    if (dtype == 0) std::memcpy(w.host.data(), host_ptr, n * sizeof(float));
    else for (size_t i = 0; i < n; ++i) w.host[i] = (float)reinterpret_cast<const double*>(host_ptr)[i];
    void* d = nullptr;
    TRY(m.dev_alloc(&d, n * sizeof(float)));
    w.dev = reinterpret_cast<float*>(d);
    PCNN_CHECK_CUDA(cudaMemcpy(w.dev, w.host.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    m.finalized = false;
    m.bn.clear();
    m.tc.clear();
    m.deconv_tc.clear();
    m.deconv_f32.clear();
    return PCNN_OK;
    ENG_GUARD_END
}

extern "C" int pcnn_finalize_weights(pcnn_handle handle, int precision) {
    PCNN_CHECK_ARG(handle, "pcnn_finalize_weights: null handle");
    ENG_GUARD_BEGIN
    eng::DeviceGuard g(handle->h.model.device);
    return eng::finalize(handle->h.model, precision, nullptr);
    ENG_GUARD_END
}

extern "C" int pcnn_set_microbatch(pcnn_handle handle, int samples) {
    PCNN_CHECK_ARG(handle && samples >= 0, "pcnn_set_microbatch: bad argument");
    handle->h.model.microbatch = samples;
    handle->h.model.need_cache.clear();
    return PCNN_OK;
}

extern "C" int pcnn_workspace_bytes(pcnn_handle handle, int B, int H, int W, size_t* bytes) {
    PCNN_CHECK_ARG(handle && bytes, "pcnn_workspace_bytes: null argument");
    ENG_GUARD_BEGIN
    eng::Model& m = handle->h.model;
    if (m.hp.present && m.db.present)
        return eng::pcnn_entry(m, handle->h.jacobi_iterations, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B, H, W, nullptr, 0, nullptr, true, bytes);
    if (m.hp.present) return eng::hpnn_entry(m, nullptr, nullptr, nullptr, B, H, W, nullptr, 0, nullptr, true, bytes);
    set_error("pcnn_workspace_bytes: use pcnn_dbcnn_workspace_bytes for a DBCNN-only handle");
    return PCNN_ERR_INVALID_ARGUMENT;
    ENG_GUARD_END
}

extern "C" int pcnn_hpnn_workspace_bytes(pcnn_handle handle, int B, int H, int W, size_t* bytes) {
    PCNN_CHECK_ARG(handle && bytes, "pcnn_hpnn_workspace_bytes: null argument");
    ENG_GUARD_BEGIN
    return eng::hpnn_entry(handle->h.model, nullptr, nullptr, nullptr, B, H, W, nullptr, 0, nullptr, true, bytes);
    ENG_GUARD_END
}

extern "C" int pcnn_dbcnn_workspace_bytes(pcnn_handle handle, int B, int n, int x_res, size_t* bytes) {
    PCNN_CHECK_ARG(handle && bytes, "pcnn_dbcnn_workspace_bytes: null argument");
    ENG_GUARD_BEGIN
    return eng::dbcnn_entry(handle->h.model, nullptr, nullptr, nullptr, B, n, x_res, nullptr, 0, nullptr, true, bytes);
    ENG_GUARD_END
}

extern "C" int pcnn_hpnn_forward(pcnn_handle handle, const float* rhs, const float* dx, float* out, int B, int H, int W,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    PCNN_CHECK_ARG(handle && rhs && dx && out, "pcnn_hpnn_forward: null argument");
    ENG_GUARD_BEGIN
    eng::DeviceGuard g(handle->h.model.device);
    return eng::hpnn_entry(handle->h.model, rhs, dx, out, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream, false, nullptr);
    ENG_GUARD_END
}

extern "C" int pcnn_dbcnn_forward(pcnn_handle handle, const float* bc, const float* dx, float* out, int B, int n, int x_res,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    PCNN_CHECK_ARG(handle && bc && dx && out, "pcnn_dbcnn_forward: null argument");
    ENG_GUARD_BEGIN
    eng::DeviceGuard g(handle->h.model.device);
    return eng::dbcnn_entry(handle->h.model, bc, dx, out, B, n, x_res, workspace, workspace_bytes, (cudaStream_t)stream, false, nullptr);
    ENG_GUARD_END
}

extern "C" int pcnn_forward(pcnn_handle handle, const float* rhs, const float* left, const float* top, const float* right,
                            const float* bottom, const float* dx, float* out, int B, int H, int W, void* workspace,
                            size_t workspace_bytes, void* stream) {
    PCNN_CHECK_ARG(handle && rhs && left && top && right && bottom && dx && out, "pcnn_forward: null argument");
    ENG_GUARD_BEGIN
    eng::DeviceGuard g(handle->h.model.device);
    return eng::pcnn_entry(handle->h.model, handle->h.jacobi_iterations, rhs, left, top, right, bottom, dx, out, B, H, W, workspace,
                           workspace_bytes, (cudaStream_t)stream, false, nullptr);
    ENG_GUARD_END
}

extern "C" int pcnn_profile_conv_begin(pcnn_handle handle, int cin, int cout, int k, int max_launches) {
    PCNN_CHECK_ARG(handle && max_launches > 0 && max_launches <= 4096, "pcnn_profile_conv_begin: bad argument");
    eng::Model& m = handle->h.model;
    eng::DeviceGuard g(m.device);
    for (auto e : m.prof_ev) cudaEventDestroy(e);
    m.prof_ev.assign((size_t)2 * max_launches, nullptr);
    for (auto& e : m.prof_ev) PCNN_CHECK_CUDA(cudaEventCreate(&e));
    m.prof_cin = cin; m.prof_cout = cout; m.prof_k = k; m.prof_used = 0; m.prof_flops = 0.0;
    return PCNN_OK;
}

extern "C" int pcnn_profile_conv_end(pcnn_handle handle, int* launches, double* avg_ms, double* flops_per_launch) {
    PCNN_CHECK_ARG(handle && launches && avg_ms && flops_per_launch, "pcnn_profile_conv_end: null argument");
    eng::Model& m = handle->h.model;
    eng::DeviceGuard g(m.device);
    double sum = 0.0;
    const int n = (int)(m.prof_used / 2);
    for (int i = 0; i < n; ++i) {
        PCNN_CHECK_CUDA(cudaEventSynchronize(m.prof_ev[2 * i + 1]));
        float ms = 0.f;
        PCNN_CHECK_CUDA(cudaEventElapsedTime(&ms, m.prof_ev[2 * i], m.prof_ev[2 * i + 1]));
        sum += ms;
    }
    *launches = n;
    *avg_ms = n ? sum / n : 0.0;
    *flops_per_launch = n ? m.prof_flops / n : 0.0;
    for (auto e : m.prof_ev) cudaEventDestroy(e);
    m.prof_ev.clear();
    m.prof_k = m.prof_cin = m.prof_cout = 0;
    m.prof_used = 0;
    return PCNN_OK;
}

// ---- host-side table builders, exported so that the CPU test-suite can pin them against the Python host's tables
// (tests/test_host.py); no GPU involved.  what: "pos" (a = n) -> float[n]; "sinh" (a = modes, b = x_res) -> float[a*b];
// "resize_idx" / "resize_w" (a = n_in, b = n_out, c = method 0 nearest / 1 bilinear / 2 bicubic) -> int32 / float [b*taps];
// "spp1" (a = n; levels = the shipped DBCNN levels are passed as `levels`, n_levels) -> int32[bins*4].
// Returns the number of elements written, or a negative status.
extern "C" long long pcnn_host_table(const char* what, int a, int b, int c, void* out, size_t out_bytes) {
    PCNN_CHECK_ARG(what && out, "pcnn_host_table: null argument");
    ENG_GUARD_BEGIN
    const std::string w(what);
    auto emit = [&](const void* src, size_t bytes, size_t count) -> long long {
        if (bytes > out_bytes) { set_error("pcnn_host_table: output buffer too small (%zu > %zu)", bytes, out_bytes); return PCNN_ERR_INVALID_ARGUMENT; }
        std::memcpy(out, src, bytes);
        return (long long)count;
    };
    if (w == "pos") { const auto v = eng::position_table(a); return emit(v.data(), v.size() * 4, v.size()); }
    if (w == "sinh") { const auto v = eng::sinh_basis(a, b); return emit(v.data(), v.size() * 4, v.size()); }
    if (w == "resize_idx") { const auto t = eng::resize_axis_table(a, b, c); return emit(t.idx.data(), t.idx.size() * 4, t.idx.size()); }
    if (w == "resize_w") { const auto t = eng::resize_axis_table(a, b, c); return emit(t.w.data(), t.w.size() * 4, t.w.size()); }
    set_error("pcnn_host_table: unknown table '%s'", what);
    return PCNN_ERR_INVALID_ARGUMENT;
    ENG_GUARD_END
}

// The separable DBCNN layer's row weights exactly as the engine uploads them: kernel [k,k,Cin,Cout] (host, fp32), M = Cin - 2
// sinh modes, grid height x_res -> fp16 image (pcnn_conv2d_tc_rowweights layout) + its power-of-two accumulator scale.
extern "C" long long pcnn_host_rowweights(const float* kernel, int k, int Cin, int Cout, int x_res, void* out_img, size_t out_bytes,
                                          float* acc_scale) {
    PCNN_CHECK_ARG(kernel && out_img && acc_scale && k >= 1 && Cin >= 3 && Cout >= 1, "pcnn_host_rowweights: bad argument");
    ENG_GUARD_BEGIN
    const int cp = pcnn_conv_tc_channel_slots(Cout, k), T = pcnn_conv_tc_rowweight_slots(Cout, k, x_res);
    PCNN_CHECK_ARG(cp > 0 && T > 0, "pcnn_host_rowweights: unsupported layer (Cout <= 32, odd k <= 15)");
    eng::Weight w;
    w.shape = {k, k, Cin, Cout};
    w.host.assign(kernel, kernel + (size_t)k * k * Cin * Cout);
    const eng::RowWeights r = eng::build_rowweights(w, Cin - 2, x_res, cp, cp == 24 ? 5 : 128 / cp, T);
    if (r.img.size() * 2 > out_bytes) { set_error("pcnn_host_rowweights: output buffer too small"); return PCNN_ERR_INVALID_ARGUMENT; }
    std::memcpy(out_img, r.img.data(), r.img.size() * 2);
    *acc_scale = r.acc_scale;
    return (long long)r.img.size();
    ENG_GUARD_END
}
