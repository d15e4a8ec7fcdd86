// pinn_kernels.cu — sm_100a kernels and the C ABI (include/pinn_b200.h) of the fused PINN fit step.
//
// One launch of step_kernel does, for every collocation point of the batch, everything the
// reference does between sampling and loss.backward() (pydens/model_torch.py:430-460):
//   * coordinates are sampled in-kernel (Philox4x32-10) or read coalesced from HBM;
//   * the flat parameter buffer is staged into shared memory with one TMA bulk copy
//     (cp.async.bulk + mbarrier) and re-laid out for broadcast 128-bit reads;
//   * one thread per point: forward jets -> ansatz -> residual program -> reverse sweep
//     (pinn_device.cuh), all per-point state resident in shared memory (or, for networks too wide
//     for that, in a global spill area);
//   * weight gradients are reduced warp -> CTA -> grid (last-arriving CTA folds the per-CTA
//     partials in a fixed order, so results are run-to-run deterministic).
// Persistent grid: one CTA per SM, every warp strides over 32-point tiles.

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <new>

#include "pinn_device.cuh"

namespace pinn {

constexpr int RS = 32;                    // row stride of per-point storage: one warp tile

struct StepArgs {
    const DevPlan* plan;
    const float* params;
    const float* points;
    uint64_t seed;
    const uint64_t* step_ptr;
    uint64_t step_val;
    uint64_t point_offset;
    long long n_points;
    float inv_n;
    float* out;
    float* residual;
    float* partials;
    unsigned int* ticket;
    float* spill;
    int n_wacc;                            // accumulator copies in smem: warps per CTA, or 1 (atomics)
    int rows_total;
};

struct FwdArgs {
    const DevPlan* plan;
    const float* params;
    const float* points;
    long long n_points;
    float* u_out;
    float* spill;
    int rows_total;
    int row_scr;
};

// ---- PTX helpers: mbarrier + 1-D TMA bulk copy (global -> shared) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// Shared-memory carve-up (in floats) common to both kernels.
struct SmemLayout {
    int plan_f, weights_f, wacc_f, bar_f, storage_f, total_f;
};
__host__ __device__ inline int align4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline SmemLayout smem_layout(int weights_floats, int n_out_floats, int n_wacc,
                                                  int storage_floats) {
    SmemLayout L;
    L.plan_f = 0;
    L.weights_f = align4((int)(sizeof(DevPlan) / 4));
    L.wacc_f = L.weights_f + align4(weights_floats);
    L.bar_f = L.wacc_f + align4(n_out_floats * n_wacc);
    L.storage_f = L.bar_f + 4;
    L.total_f = L.storage_f + align4(storage_floats);
    return L;
}

// Stage plan + parameters into shared memory.  Returns with __syncthreads() done.
__device__ __forceinline__ void stage_plan_and_weights(float* smem, const SmemLayout& SL, const DevPlan* gplan,
                                                       const float* params) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // plan copy
    {
        const uint4* src = reinterpret_cast<const uint4*>(gplan);
        uint4* dst = reinterpret_cast<uint4*>(smem + SL.plan_f);
        for (int i = tid; i < (int)(sizeof(DevPlan) / 16); i += nt) dst[i] = src[i];
    }
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SL.bar_f);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    const DevPlan& P = *reinterpret_cast<const DevPlan*>(smem + SL.plan_f);
    float* stage = smem + SL.storage_f;           // parameters land here first
    if (tid == 0) {
        uint32_t bytes = (uint32_t)P.n_params * 4u;
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(stage, params, bytes, bar);
    }
    float* sw = smem + SL.weights_f;
    for (int i = tid; i < P.weights_floats; i += nt) sw[i] = 0.0f;
    mbar_wait(bar, 0);
    __syncthreads();
    for (int l = 0; l < P.n_layers; ++l) {
        const DevLayer& L = P.layer[l];
        const int n = L.n_in * L.n_out;
        for (int i = tid; i < n; i += nt) {
            int j = i / L.n_in, k = i - j * L.n_in;
            float w = stage[L.w_off + i];
            sw[L.wt_s + k * L.n_out_p4 + j] = w;
            sw[L.w_s + j * L.n_in_p8 + k] = w;
        }
        for (int j = tid; j < L.n_out; j += nt) sw[L.b_s + j] = stage[L.b_off + j];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------
// The fit-step kernel.
// ---------------------------------------------------------------------------------------------------
template <int NF, int NS, bool GMEM, int MAXT, int JF>
__global__ void __launch_bounds__(MAXT, 1) step_kernel(const StepArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    const DevPlan& P = *reinterpret_cast<const DevPlan*>(smem);
    // sizes needed for the carve-up come straight from the global plan (uniform loads)
    const int n_out_floats = a.plan->n_params + 4;
    const SmemLayout SL = smem_layout(a.plan->weights_floats, n_out_floats, a.n_wacc,
                                      GMEM ? a.plan->n_params : max(a.plan->n_params, a.rows_total * RS * nwarps));
    stage_plan_and_weights(smem, SL, a.plan, a.params);
    const float* sw = smem + SL.weights_f;
    float* wacc_all = smem + SL.wacc_f;
    for (int i = tid; i < n_out_floats * a.n_wacc; i += blockDim.x) wacc_all[i] = 0.0f;
    __syncthreads();

    GradSink sink;
    sink.atomic = (a.n_wacc == 1 && nwarps > 1);
    sink.wacc = wacc_all + (a.n_wacc == 1 ? 0 : warp * n_out_floats);

    const long long gw = (long long)blockIdx.x * nwarps + warp;         // global warp id
    const long long total_warps = (long long)gridDim.x * nwarps;
    float* st = (GMEM ? a.spill + (size_t)gw * a.rows_total * RS : smem + SL.storage_f + (size_t)warp * a.rows_total * RS) + lane;

    const uint64_t step = a.step_ptr ? *a.step_ptr : a.step_val;
    const long long n_tiles = (a.n_points + 31) / 32;
    PointPartials<NF, NS> part;
    part.loss = 0.0f; part.sbar = 0.0f;
#pragma unroll
    for (int i = 0; i < PINN_MAX_VARS; ++i) part.vbar[i] = 0.0f;

    for (long long tile = gw; tile < n_tiles; tile += total_warps) {
        const long long pl = tile * 32 + lane;
        const bool valid = pl < a.n_points;
        const long long pe = valid ? pl : a.n_points - 1;     // masked lanes replay the last point
        if (a.points) {
            const float* src = a.points + (size_t)pe * P.total;
            for (int k = 0; k < P.total; ++k) st[k * RS] = __ldg(src + k);
        } else {
            const uint64_t gidx = a.point_offset + (uint64_t)pe;
            const uint32_t c3 = (uint32_t)(((step >> 32) & 0xffffu) << 16);
            Philox4 b0 = philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)step, c3,
                                       (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            Philox4 b1 = b0;
            if (P.total > 4)
                b1 = philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)step, c3 | 1u,
                                   (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            for (int k = 0; k < P.total; ++k) st[k * RS] = sample_column(P.cols[k], k, gidx, step, a.seed, b0, b1);
        }
        float r = point_step<NF, NS, JF>(P, sw, a.params, st, RS, valid, a.inv_n, sink, part);
        if (a.residual && valid) a.residual[pl] = r;
    }

    // per-thread scalars -> accumulator
    {
        float v = warp_sum(part.loss);
        if (lane == 0) sink.add(P.n_params, v);
        v = warp_sum(part.sbar);
        if (lane == 0) sink.add(P.log_scale_off, v);
#pragma unroll
        for (int i = 0; i < PINN_MAX_VARS; ++i) {
            if (i < P.n_vars) {
                float t = warp_sum(part.vbar[i]);
                if (lane == 0) sink.add(P.var_off[i], t);
            }
        }
    }
    __syncthreads();

    // CTA partial -> global, then the last CTA folds all partials in block order
    float* mine = a.partials + (size_t)blockIdx.x * n_out_floats;
    for (int i = tid; i < n_out_floats; i += blockDim.x) {
        float s = 0.0f;
        for (int w = 0; w < a.n_wacc; ++w) s += wacc_all[w * n_out_floats + i];
        mine[i] = s;
    }
    __threadfence();
    __syncthreads();
    __shared__ unsigned int s_last;
    if (tid == 0) {
        unsigned int t = atomicAdd(a.ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int i = tid; i < n_out_floats; i += blockDim.x) {
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
            int b = 0;
            const float* src = a.partials + i;
            for (; b + 3 < (int)gridDim.x; b += 4) {
                s0 += __ldcg(src + (size_t)(b + 0) * n_out_floats);
                s1 += __ldcg(src + (size_t)(b + 1) * n_out_floats);
                s2 += __ldcg(src + (size_t)(b + 2) * n_out_floats);
                s3 += __ldcg(src + (size_t)(b + 3) * n_out_floats);
            }
            for (; b < (int)gridDim.x; ++b) s0 += __ldcg(src + (size_t)b * n_out_floats);
            a.out[i] = (s0 + s1) + (s2 + s3);
        }
        if (tid == 0) *a.ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------
// Forward-only kernel (predict): u for explicit points.
// ---------------------------------------------------------------------------------------------------
template <bool GMEM>
__global__ void __launch_bounds__(256, 1) forward_kernel(const FwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const DevPlan& P = *reinterpret_cast<const DevPlan*>(smem);
    const SmemLayout SL = smem_layout(a.plan->weights_floats, 0, 0,
                                      GMEM ? a.plan->n_params : max(a.plan->n_params, a.rows_total * RS * nwarps));
    stage_plan_and_weights(smem, SL, a.plan, a.params);
    const float* sw = smem + SL.weights_f;
    const long long gw = (long long)blockIdx.x * nwarps + warp;
    const long long total_warps = (long long)gridDim.x * nwarps;
    float* st = (GMEM ? a.spill + (size_t)gw * a.rows_total * RS : smem + SL.storage_f + (size_t)warp * a.rows_total * RS) + lane;
    const long long n_tiles = (a.n_points + 31) / 32;
    for (long long tile = gw; tile < n_tiles; tile += total_warps) {
        const long long pl = tile * 32 + lane;
        const bool valid = pl < a.n_points;
        const long long pe = valid ? pl : a.n_points - 1;
        const float* src = a.points + (size_t)pe * P.total;
        for (int k = 0; k < P.total; ++k) st[k * RS] = __ldg(src + k);
        float u = point_forward<16>(P, sw, a.params, st, RS, a.row_scr);
        if (valid) a.u_out[pl] = u;
    }
}

// Points the in-kernel sampler produces (for tests / replay).
__global__ void sample_kernel(const DevPlan* plan, uint64_t seed, const uint64_t* step_ptr, uint64_t step_val,
                              uint64_t point_offset, long long n_points, float* out) {
    const uint64_t step = step_ptr ? *step_ptr : step_val;
    const int total = plan->total;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_points;
         p += (long long)gridDim.x * blockDim.x) {
        const uint64_t gidx = point_offset + (uint64_t)p;
        const uint32_t c3 = (uint32_t)(((step >> 32) & 0xffffu) << 16);
        Philox4 b0 = philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)step, c3, (uint32_t)seed,
                                   (uint32_t)(seed >> 32));
        Philox4 b1 = b0;
        if (total > 4)
            b1 = philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)step, c3 | 1u, (uint32_t)seed,
                               (uint32_t)(seed >> 32));
        for (int k = 0; k < total; ++k)
            out[(size_t)p * total + k] = sample_column(plan->cols[k], k, gidx, step, seed, b0, b1);
    }
}

__global__ void record_loss_kernel(const float* out, int loss_idx, float* ring, long long ring_len,
                                   unsigned long long* step) {
    unsigned long long s = *step;
    ring[s % (unsigned long long)ring_len] = out[loss_idx];
    *step = s + 1ull;
}

__global__ void set_cols_kernel(DevPlan* plan, PinnColumn c0, PinnColumn c1, PinnColumn c2, PinnColumn c3,
                                PinnColumn c4, PinnColumn c5, PinnColumn c6, PinnColumn c7) {
    plan->cols[0] = c0; plan->cols[1] = c1; plan->cols[2] = c2; plan->cols[3] = c3;
    plan->cols[4] = c4; plan->cols[5] = c5; plan->cols[6] = c6; plan->cols[7] = c7;
}

}  // namespace pinn

// ===================================================================================================
// Host side: plan construction, variant dispatch, C ABI.
// ===================================================================================================
using namespace pinn;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) return fail(PINN_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

typedef void (*StepKernelFn)(const StepArgs);

struct Variant {
    int nf, ns;
    StepKernelFn smem_fn, gmem_fn;
    int maxt;
};

template <int NF, int NS>
struct VariantCfg {
    static constexpr int C = 1 + NF + NS;
    static constexpr int MAXT = (C <= 3) ? 512 : 256;
    static constexpr int JF = 16;
};

template <int NF, int NS>
static Variant make_variant() {
    using Cfg = VariantCfg<NF, NS>;
    Variant v;
    v.nf = NF; v.ns = NS;
    v.smem_fn = step_kernel<NF, NS, false, Cfg::MAXT, Cfg::JF>;
    v.gmem_fn = step_kernel<NF, NS, true, Cfg::MAXT, Cfg::JF>;
    v.maxt = Cfg::MAXT;
    return v;
}

static const Variant* find_variant(int nf, int ns) {
    static const Variant table[] = {
        make_variant<0, 0>(),
        make_variant<1, 0>(), make_variant<1, 1>(),
        make_variant<2, 0>(), make_variant<2, 1>(), make_variant<2, 2>(),
        make_variant<3, 0>(), make_variant<3, 1>(), make_variant<3, 2>(), make_variant<3, 3>(),
        make_variant<4, 0>(), make_variant<4, 1>(), make_variant<4, 2>(), make_variant<4, 3>(), make_variant<4, 4>(),
    };
    for (const Variant& v : table)
        if (v.nf == nf && v.ns == ns) return &v;
    return nullptr;
}

struct PinnPlan {
    PinnSpec spec;
    DevPlan h;
    DevPlan* d;
    int device;
    const Variant* var;
    int sm_count;
    int smem_optin;
    // step kernel launch config
    bool gmem;
    int threads, n_wacc, smem_bytes, regs;
    // forward kernel launch config
    bool fwd_gmem;
    int fwd_threads, fwd_smem_bytes, fwd_rows, fwd_row_scr;
    PinnColumn cur_cols[PINN_MAX_DIMS];
    bool cols_set;
};

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

extern "C" const char* pinn_last_error(void) { return g_err; }
extern "C" int pinn_abi_version(void) { return PINN_ABI_VERSION; }

static int validate_prog(const PinnInstr* prog, int n, int n_slots, int total, int n_vars, const char* name) {
    if (n < 0 || n > PINN_MAX_PROG) return fail(PINN_E_INVALID, "%s: %d instructions (max %d)", name, n, PINN_MAX_PROG);
    for (int i = 0; i < n; ++i) {
        const PinnInstr& in = prog[i];
        if (in.op >= PINN_OP_COUNT_) return fail(PINN_E_INVALID, "%s[%d]: bad opcode %d", name, i, in.op);
        if (in.dst >= n_slots) return fail(PINN_E_INVALID, "%s[%d]: dst slot %d >= n_slots %d", name, i, in.dst, n_slots);
        if (in.op == PINN_OP_COORD && in.a >= total) return fail(PINN_E_INVALID, "%s[%d]: coord %d", name, i, in.a);
        if (in.op == PINN_OP_VAR && in.a >= n_vars) return fail(PINN_E_INVALID, "%s[%d]: var %d", name, i, in.a);
        if (in.op >= PINN_OP_ADD) {
            if (in.a >= n_slots) return fail(PINN_E_INVALID, "%s[%d]: src slot", name, i);
            bool binary = in.op == PINN_OP_ADD || in.op == PINN_OP_SUB || in.op == PINN_OP_MUL ||
                          in.op == PINN_OP_DIV || in.op == PINN_OP_POW;
            if (binary && in.b >= n_slots) return fail(PINN_E_INVALID, "%s[%d]: src slot", name, i);
        }
    }
    return PINN_OK;
}

extern "C" int pinn_plan_create(const PinnSpec* s, int device, PinnPlan** out) {
    if (!s || !out) return fail(PINN_E_INVALID, "null argument");
    if (s->abi_version != PINN_ABI_VERSION)
        return fail(PINN_E_INVALID, "spec abi_version %d != library %d", s->abi_version, PINN_ABI_VERSION);
    const int Ln = s->n_layers;
    if (Ln < 1 || Ln > PINN_MAX_LAYERS) return fail(PINN_E_INVALID, "n_layers %d", Ln);
    const int total = s->ndims + s->nparams;
    if (s->ndims < 1 || s->nparams < 0 || total > PINN_MAX_DIMS) return fail(PINN_E_INVALID, "ndims/nparams");
    if (s->widths[0] != total) return fail(PINN_E_INVALID, "widths[0]=%d != ndims+nparams=%d", s->widths[0], total);
    if (s->widths[Ln] != 1) return fail(PINN_E_UNSUPPORTED, "output width %d (must be 1)", s->widths[Ln]);
    if (s->act[Ln - 1] != PINN_ACT_NONE) return fail(PINN_E_UNSUPPORTED, "activation after the last layer");
    if (s->n_params <= 0 || (s->n_params & 3)) return fail(PINN_E_INVALID, "n_params must be a positive multiple of 4");
    if (s->nf < 0 || s->nf > PINN_MAX_DIRS || s->ns < 0 || s->ns > s->nf) return fail(PINN_E_INVALID, "jet set nf=%d ns=%d", s->nf, s->ns);
    if (s->n_vars < 0 || s->n_vars > PINN_MAX_VARS) return fail(PINN_E_INVALID, "n_vars");
    const int C = 1 + s->nf + s->ns;
    for (int d = 0; d < s->nf; ++d)
        if (s->dir_col[d] < 0 || s->dir_col[d] >= total) return fail(PINN_E_INVALID, "dir_col[%d]", d);
    if (s->n_slots < C || s->n_slots > PINN_MAX_SLOTS) return fail(PINN_E_INVALID, "n_slots %d", s->n_slots);
    int rc;
    if ((rc = validate_prog(s->eq_prog, s->n_eq, s->n_slots, total, s->n_vars, "eq_prog"))) return rc;
    if ((rc = validate_prog(s->ic_prog, s->n_ic, s->n_slots, total, 0, "ic_prog"))) return rc;
    for (int i = 0; i < 1 + C + s->n_vars; ++i)
        if (s->eq_out[i] < 0 || s->eq_out[i] >= s->n_slots) return fail(PINN_E_INVALID, "eq_out[%d]", i);
    if (s->has_ic)
        for (int c = 0; c < C; ++c)
            if (s->ic_out[c] < 0 || s->ic_out[c] >= s->n_slots) return fail(PINN_E_INVALID, "ic_out[%d]", c);
    const Variant* var = find_variant(s->nf, s->ns);
    if (!var) return fail(PINN_E_UNSUPPORTED, "no kernel variant for nf=%d ns=%d", s->nf, s->ns);

    PinnPlan* p = new (std::nothrow) PinnPlan();
    if (!p) return fail(PINN_E_INVALID, "out of memory");
    p->spec = *s;
    p->device = device;
    p->var = var;
    p->d = nullptr;
    p->cols_set = false;
    DevPlan& h = p->h;
    memset(&h, 0, sizeof(h));
    h.n_layers = Ln; h.total = total; h.ndims = s->ndims; h.nparams = s->nparams;
    h.has_bc = s->has_bc ? 1 : 0; h.has_ic = s->has_ic ? 1 : 0;
    h.nsp = s->has_ic ? s->ndims - 1 : s->ndims;
    h.nf = s->nf; h.ns = s->ns; h.n_params = s->n_params; h.log_scale_off = s->log_scale_off;
    h.n_vars = s->n_vars; h.n_eq = s->n_eq; h.n_ic = s->has_ic ? s->n_ic : 0; h.n_slots = s->n_slots;
    h.bc = s->bc_value;
    h.t0 = s->dom_lo[s->ndims - 1];
    for (int d = 0; d < PINN_MAX_DIRS; ++d) h.dir_col[d] = d < s->nf ? s->dir_col[d] : 0;
    for (int i = 0; i < PINN_MAX_VARS; ++i) h.var_off[i] = i < s->n_vars ? s->var_off[i] : 0;
    for (int i = 0; i < PINN_MAX_DIMS; ++i) {
        h.lo[i] = s->dom_lo[i]; h.hi[i] = s->dom_hi[i];
        float w = s->dom_hi[i] - s->dom_lo[i];
        h.inv_w2[i] = (i < s->ndims && w != 0.0f) ? 1.0f / (w * w) : 0.0f;
        h.cols[i].kind = PINN_COL_UNIFORM; h.cols[i].a = 0.0f; h.cols[i].b = 1.0f;
    }
    memcpy(h.eq_out, s->eq_out, sizeof(h.eq_out));
    memcpy(h.ic_out, s->ic_out, sizeof(h.ic_out));
    memcpy(h.eq, s->eq_prog, sizeof(h.eq));
    memcpy(h.ic, s->ic_prog, sizeof(h.ic));
    if (s->log_scale_off < 0 || s->log_scale_off >= s->n_params) { delete p; return fail(PINN_E_INVALID, "log_scale_off"); }

    int sw = 0, units = 0;
    long long macs = 0;
    for (int l = 0; l < Ln; ++l) {
        DevLayer& L = h.layer[l];
        L.n_in = s->widths[l]; L.n_out = s->widths[l + 1]; L.act = s->act[l];
        if (L.n_in < 1 || L.n_out < 1) { delete p; return fail(PINN_E_INVALID, "layer %d width", l); }
        if (L.act < 0 || L.act > PINN_ACT_SIN) { delete p; return fail(PINN_E_INVALID, "layer %d activation", l); }
        L.w_off = s->w_off[l]; L.b_off = s->b_off[l];
        if (L.w_off < 0 || L.w_off + L.n_in * L.n_out > s->n_params || L.b_off < 0 || L.b_off + L.n_out > s->n_params) {
            delete p; return fail(PINN_E_INVALID, "layer %d offsets out of range", l);
        }
        L.n_out_p4 = round_up(L.n_out, 4);
        L.n_in_p8 = round_up(L.n_in, 8);
        L.wt_s = sw; sw += L.n_in * L.n_out_p4;
        L.w_s = sw;  sw += L.n_out_p4 * L.n_in_p8;
        L.b_s = sw;  sw += round_up(L.n_out_p4, 16);     // forward blocks read bias[j0 .. j0+15]
        L.unit_base = units; units += L.n_out;
        macs += (long long)L.n_in * L.n_out;
    }
    // forward weight rows are read 16 floats at a time from j0: pad the tail
    sw += 16;
    h.weights_floats = round_up(sw, 4);
    h.n_units = units;
    h.row_units = total;
    h.row_scr = total + units * C;
    h.rows_total = h.row_scr + s->n_slots;
    p->fwd_row_scr = total + units;
    p->fwd_rows = p->fwd_row_scr + s->n_slots;

    // ---- device queries and launch configuration ----
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete p; return fail(PINN_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e)); }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete p; return fail(PINN_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    if (prop.major != 10) { delete p; return fail(PINN_E_UNSUPPORTED, "device sm_%d%d: this library is built for sm_100a only", prop.major, prop.minor); }
    p->sm_count = prop.multiProcessorCount;
    p->smem_optin = (int)prop.sharedMemPerBlockOptin;

    const int n_out_floats = s->n_params + 4;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, (const void*)var->smem_fn);
    if (e != cudaSuccess) { delete p; return fail(PINN_E_CUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(e)); }
    p->regs = fa.numRegs;
    int max_warps_regs = (65536 / (fa.numRegs * 32));
    int max_warps = var->maxt / 32;
    if (max_warps_regs < max_warps) max_warps = max_warps_regs;
    // shared-memory resident activations: largest warp count that fits (per-warp accumulators
    // when they fit too, else one shared accumulator with atomics)
    const int budget = p->smem_optin - (int)fa.sharedSizeBytes - 64;
    int best_nw = 0, best_nwacc = 0;
    for (int nw = max_warps; nw >= 1 && !best_nw; --nw) {
        for (int pass = 0; pass < 2 && !best_nw; ++pass) {
            int nwacc = pass == 0 ? nw : 1;
            SmemLayout SL = smem_layout(h.weights_floats, n_out_floats, nwacc,
                                        h.rows_total * RS * nw > h.n_params ? h.rows_total * RS * nw : h.n_params);
            if (SL.total_f * 4 <= budget) { best_nw = nw; best_nwacc = nwacc; }
        }
    }
    if (best_nw >= 4 || (best_nw >= 2 && max_warps <= 4)) {
        p->gmem = false; p->threads = best_nw * 32; p->n_wacc = best_nwacc;
        SmemLayout SL = smem_layout(h.weights_floats, n_out_floats, best_nwacc,
                                    h.rows_total * RS * best_nw > h.n_params ? h.rows_total * RS * best_nw : h.n_params);
        p->smem_bytes = SL.total_f * 4;
    } else {
        // activations spill to a global workspace; accumulators: per warp if they fit, else shared
        e = cudaFuncGetAttributes(&fa, (const void*)var->gmem_fn);
        if (e != cudaSuccess) { delete p; return fail(PINN_E_CUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(e)); }
        p->regs = fa.numRegs;
        int nw = var->maxt / 32;
        if (65536 / (fa.numRegs * 32) < nw) nw = 65536 / (fa.numRegs * 32);
        int nwacc = nw;
        SmemLayout SL = smem_layout(h.weights_floats, n_out_floats, nwacc, h.n_params);
        if (SL.total_f * 4 > budget) { nwacc = 1; SL = smem_layout(h.weights_floats, n_out_floats, 1, h.n_params); }
        if (SL.total_f * 4 > budget) { delete p; return fail(PINN_E_UNSUPPORTED, "network too large: %d B of weights do not fit shared memory", SL.total_f * 4); }
        p->gmem = true; p->threads = nw * 32; p->n_wacc = nwacc; p->smem_bytes = SL.total_f * 4;
    }
    e = cudaFuncSetAttribute((const void*)(p->gmem ? var->gmem_fn : var->smem_fn),
                             cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes);
    if (e != cudaSuccess) { delete p; return fail(PINN_E_CUDA, "cudaFuncSetAttribute(smem=%d): %s", p->smem_bytes, cudaGetErrorString(e)); }

    // forward kernel config
    {
        int nw = 8;
        SmemLayout SL = smem_layout(h.weights_floats, 0, 0, p->fwd_rows * RS * nw > h.n_params ? p->fwd_rows * RS * nw : h.n_params);
        if (SL.total_f * 4 <= p->smem_optin - 64) { p->fwd_gmem = false; }
        else { p->fwd_gmem = true; SL = smem_layout(h.weights_floats, 0, 0, h.n_params); }
        p->fwd_threads = nw * 32; p->fwd_smem_bytes = SL.total_f * 4;
        if (p->fwd_smem_bytes > p->smem_optin - 64) { delete p; return fail(PINN_E_UNSUPPORTED, "network too large for shared memory"); }
        e = cudaFuncSetAttribute((const void*)(p->fwd_gmem ? forward_kernel<true> : forward_kernel<false>),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, p->fwd_smem_bytes);
        if (e != cudaSuccess) { delete p; return fail(PINN_E_CUDA, "cudaFuncSetAttribute(fwd): %s", cudaGetErrorString(e)); }
    }

    e = cudaMalloc(&p->d, sizeof(DevPlan));
    if (e != cudaSuccess) { delete p; return fail(PINN_E_CUDA, "cudaMalloc(plan): %s", cudaGetErrorString(e)); }
    e = cudaMemcpy(p->d, &h, sizeof(DevPlan), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p->d); delete p; return fail(PINN_E_CUDA, "cudaMemcpy(plan): %s", cudaGetErrorString(e)); }
    for (int i = 0; i < PINN_MAX_DIMS; ++i) p->cur_cols[i] = h.cols[i];
    p->cols_set = true;
    (void)macs;
    *out = p;
    return PINN_OK;
}

extern "C" int pinn_plan_destroy(PinnPlan* p) {
    if (!p) return PINN_OK;
    if (p->d) cudaFree(p->d);
    delete p;
    return PINN_OK;
}

static int grid_for(const PinnPlan* p, long long n_points, int threads) {
    long long tiles = (n_points + 31) / 32;
    long long ctas = (tiles + threads / 32 - 1) / (threads / 32);
    if (ctas > p->sm_count) ctas = p->sm_count;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

// workspace layout: [ticket: 256 B][partials: sm_count x n_out_floats][spill]
static size_t ws_partials_off() { return 256; }
static size_t ws_spill_off(const PinnPlan* p) {
    size_t b = ws_partials_off() + (size_t)p->sm_count * (p->h.n_params + 4) * sizeof(float);
    return (b + 255) & ~(size_t)255;
}

extern "C" size_t pinn_workspace_bytes(const PinnPlan* p, int64_t n_points) {
    if (!p) return 0;
    size_t b = ws_spill_off(p);
    size_t spill = 0;
    if (p->gmem) spill = (size_t)p->sm_count * (p->threads / 32) * p->h.rows_total * RS * sizeof(float);
    if (p->fwd_gmem) {
        size_t f = (size_t)p->sm_count * (p->fwd_threads / 32) * p->fwd_rows * RS * sizeof(float);
        if (f > spill) spill = f;
    }
    (void)n_points;
    return b + spill;
}

extern "C" int pinn_out_floats(const PinnPlan* p) { return p ? p->h.n_params + 4 : 0; }

static bool cols_equal(const PinnColumn* a, const PinnColumn* b, int n) {
    for (int i = 0; i < n; ++i)
        if (a[i].kind != b[i].kind || a[i].a != b[i].a || a[i].b != b[i].b) return false;
    return true;
}

static int apply_cols(PinnPlan* p, const PinnColumn* cols, cudaStream_t st) {
    PinnColumn want[PINN_MAX_DIMS];
    for (int i = 0; i < PINN_MAX_DIMS; ++i) {
        if (cols && i < p->h.total) want[i] = cols[i];
        else { want[i].kind = PINN_COL_UNIFORM; want[i].a = 0.0f; want[i].b = 1.0f; }
        if (want[i].kind < 0 || want[i].kind > PINN_COL_CONST) return fail(PINN_E_INVALID, "column %d kind %d", i, want[i].kind);
    }
    if (p->cols_set && cols_equal(want, p->cur_cols, PINN_MAX_DIMS)) return PINN_OK;
    set_cols_kernel<<<1, 1, 0, st>>>(p->d, want[0], want[1], want[2], want[3], want[4], want[5], want[6], want[7]);
    CUDA_TRY(cudaGetLastError());
    for (int i = 0; i < PINN_MAX_DIMS; ++i) p->cur_cols[i] = want[i];
    p->cols_set = true;
    return PINN_OK;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

extern "C" int pinn_step(const PinnPlan* cp, const float* params, const float* points, const PinnColumn* cols,
                         uint64_t seed, const uint64_t* step_counter, uint64_t step_value, uint64_t point_offset,
                         int64_t n_points, float inv_global_n, float* grads_and_loss, float* residual_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
    PinnPlan* p = const_cast<PinnPlan*>(cp);
    if (!p || !params || !grads_and_loss || !workspace) return fail(PINN_E_INVALID, "null argument");
    if (n_points <= 0) return fail(PINN_E_INVALID, "n_points must be positive");
    if (!aligned16(params) || !aligned16(grads_and_loss) || !aligned16(workspace) || (points && !aligned16(points)))
        return fail(PINN_E_ALIGN, "device pointers must be 16-byte aligned");
    if (workspace_bytes < pinn_workspace_bytes(p, n_points))
        return fail(PINN_E_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, pinn_workspace_bytes(p, n_points));
    cudaStream_t st = (cudaStream_t)stream;
    if (!points) { int rc = apply_cols(p, cols, st); if (rc) return rc; }
    StepArgs a;
    a.plan = p->d; a.params = params; a.points = points; a.seed = seed;
    a.step_ptr = step_counter; a.step_val = step_value; a.point_offset = point_offset;
    a.n_points = n_points; a.inv_n = inv_global_n; a.out = grads_and_loss; a.residual = residual_out;
    a.ticket = reinterpret_cast<unsigned int*>(workspace);
    a.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ws_partials_off());
    a.spill = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ws_spill_off(p));
    a.n_wacc = p->n_wacc; a.rows_total = p->h.rows_total;
    const int grid = grid_for(p, n_points, p->threads);
    StepKernelFn fn = p->gmem ? p->var->gmem_fn : p->var->smem_fn;
    fn<<<grid, p->threads, p->smem_bytes, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return PINN_OK;
}

extern "C" int pinn_forward(const PinnPlan* p, const float* params, const float* points, int64_t n_points,
                            float* u_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!p || !params || !points || !u_out || !workspace) return fail(PINN_E_INVALID, "null argument");
    if (n_points <= 0) return fail(PINN_E_INVALID, "n_points must be positive");
    if (!aligned16(params) || !aligned16(points) || !aligned16(workspace)) return fail(PINN_E_ALIGN, "device pointers must be 16-byte aligned");
    if (workspace_bytes < pinn_workspace_bytes(p, n_points)) return fail(PINN_E_WORKSPACE, "workspace too small");
    FwdArgs a;
    a.plan = p->d; a.params = params; a.points = points; a.n_points = n_points; a.u_out = u_out;
    a.spill = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ws_spill_off(p));
    a.rows_total = p->fwd_rows; a.row_scr = p->fwd_row_scr;
    const int grid = grid_for(p, n_points, p->fwd_threads);
    cudaStream_t st = (cudaStream_t)stream;
    if (p->fwd_gmem) forward_kernel<true><<<grid, p->fwd_threads, p->fwd_smem_bytes, st>>>(a);
    else             forward_kernel<false><<<grid, p->fwd_threads, p->fwd_smem_bytes, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return PINN_OK;
}

extern "C" int pinn_sample(const PinnPlan* cp, const PinnColumn* cols, uint64_t seed, const uint64_t* step_counter,
                           uint64_t step_value, uint64_t point_offset, int64_t n_points, float* points_out,
                           void* stream) {
    PinnPlan* p = const_cast<PinnPlan*>(cp);
    if (!p || !points_out) return fail(PINN_E_INVALID, "null argument");
    if (n_points <= 0) return fail(PINN_E_INVALID, "n_points must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = apply_cols(p, cols, st);
    if (rc) return rc;
    long long blocks = (n_points + 255) / 256;
    if (blocks > 4 * p->sm_count) blocks = 4 * p->sm_count;
    sample_kernel<<<(int)blocks, 256, 0, st>>>(p->d, seed, step_counter, step_value, point_offset, n_points, points_out);
    CUDA_TRY(cudaGetLastError());
    return PINN_OK;
}

extern "C" int pinn_record_loss(const PinnPlan* p, const float* grads_and_loss, float* losses_ring, int64_t ring_len,
                                uint64_t* step_counter, void* stream) {
    if (!p || !grads_and_loss || !losses_ring || !step_counter || ring_len <= 0) return fail(PINN_E_INVALID, "bad argument");
    record_loss_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(grads_and_loss, p->h.n_params, losses_ring, ring_len,
                                                         reinterpret_cast<unsigned long long*>(step_counter));
    CUDA_TRY(cudaGetLastError());
    return PINN_OK;
}

extern "C" int pinn_plan_info(const PinnPlan* p, PinnPlanInfo* info) {
    if (!p || !info) return fail(PINN_E_INVALID, "null argument");
    const int C = 1 + p->h.nf + p->h.ns;
    long long macs = 0;
    for (int l = 0; l < p->h.n_layers; ++l) macs += (long long)p->h.layer[l].n_in * p->h.layer[l].n_out;
    info->nf = p->h.nf; info->ns = p->h.ns; info->channels = C;
    info->threads_per_cta = p->threads; info->ctas_per_sm = 1;
    info->activations_in_smem = p->gmem ? 0 : 1;
    info->smem_bytes = p->smem_bytes; info->regs_per_thread = p->regs; info->sm_count = p->sm_count;
    info->rows_per_point = p->h.rows_total;
    info->flops_per_point = 6ll * C * macs;
    info->bytes_per_point = 4 * p->h.total;
    return PINN_OK;
}
