// pinn_step_kernel.cuh — the fused fit-step kernel template and its shared-memory staging helpers.
// Included by pinn_kernels.cu (host API, forward/sampling kernels) and by the per-NF translation
// units pinn_variants_nf*.cu, which only instantiate step_kernel so the build can run in parallel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "pinn_device.cuh"

#ifndef PINN_COMM_MAX_RANKS
#define PINN_COMM_MAX_RANKS 8
#endif

namespace pinn {

constexpr int RS = 32;                    // row stride of per-point storage: one warp tile

struct StepArgs {
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
    // in-kernel all-reduce over NVLink peer memory (comm_world == 0: off)
    char* comm_peers[PINN_COMM_MAX_RANKS]; // exchange buffer of every rank, mapped through CUDA IPC
    int comm_rank, comm_world;
    long long comm_timeout;                // clock64 ticks a rank waits for its peers before poisoning the step
};

// Layout of one rank's exchange buffer (pinn_comm_create): epoch counter, arrival flags, slots.
constexpr int PINN_COMM_FLAGS_OFF = 256;
constexpr int PINN_COMM_SLOTS_OFF = 1024;
__host__ __device__ inline size_t comm_slot_floats(int n_out_floats) { return (size_t)((n_out_floats + 31) & ~31); }
__host__ __device__ inline size_t comm_bytes(int n_out_floats, int world) {
    return PINN_COMM_SLOTS_OFF + 2 * (size_t)world * comm_slot_floats(n_out_floats) * sizeof(float);
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}


struct FwdArgs {
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
    int weights_f, wacc_f, bar_f, storage_f, total_f;
};
__host__ __device__ inline int align4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline SmemLayout smem_layout(int weights_floats, int n_out_floats, int n_wacc,
                                                  int storage_floats) {
    SmemLayout L;
    L.weights_f = 0;
    L.wacc_f = L.weights_f + align4(weights_floats);
    L.bar_f = L.wacc_f + align4(n_out_floats * n_wacc);
    L.storage_f = L.bar_f + 4;
    L.total_f = L.storage_f + align4(storage_floats);
    return L;
}

// Stage the parameters into shared memory: one TMA bulk copy of the flat buffer, then a re-layout
// into the padded forward / reverse weight matrices.  Returns with __syncthreads() done.
// The plan itself is a __grid_constant__ kernel parameter: every loop bound and program word is read
// through the constant bank, i.e. provably warp-uniform.
__device__ __forceinline__ void stage_weights(float* smem, const SmemLayout& SL, const DevPlan& P,
                                              const float* params) {
    const int tid = threadIdx.x, nt = blockDim.x;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SL.bar_f);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
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
// Grid tail shared by every step kernel: the CTA has written its partial [n_out_floats] to
// a.partials[blockIdx.x]; the last CTA to arrive (ticket) folds all partials in block order and — in
// data-parallel runs — exchanges the folded vector with the peers over NVLink.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void finish_grid(const StepArgs& a, const int n_out_floats) {
    const int tid = threadIdx.x;
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
        // --- data-parallel runs: the all-reduce of [grads | loss] happens right here, over NVLink peer
        // memory.  This rank's folded vector is stored into its slot in EVERY rank's exchange buffer, an
        // arrival flag follows (release, system scope), and once all flags of this epoch are in, every
        // rank sums the slots in rank order — identical bits everywhere, no second kernel, no NCCL call.
        // Slots and flags are double-buffered by epoch parity: a rank can be at most one step ahead.
        __shared__ unsigned int s_epoch;
        __shared__ int s_timeout;
        if (tid == 0) s_timeout = 0;
        float* slot_base = nullptr;
        unsigned int epoch = 0;
        int par = 0;
        const size_t slot_f = comm_slot_floats(n_out_floats);
        if (a.comm_world > 1) {
            if (tid == 0) {
                unsigned int* ep = reinterpret_cast<unsigned int*>(a.comm_peers[a.comm_rank]);
                s_epoch = *ep + 1u;
                *ep = s_epoch;
            }
            __syncthreads();
            epoch = s_epoch;
            par = (int)(epoch & 1u);
            slot_base = reinterpret_cast<float*>(a.comm_peers[a.comm_rank] + PINN_COMM_SLOTS_OFF);
        }
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
            const float total = (s0 + s1) + (s2 + s3);
            if (a.comm_world > 1) {
                for (int r = 0; r < a.comm_world; ++r) {       // peer stores over NVLink
                    float* dst = reinterpret_cast<float*>(a.comm_peers[r] + PINN_COMM_SLOTS_OFF) +
                                 ((size_t)par * a.comm_world + a.comm_rank) * slot_f;
                    dst[i] = total;
                }
            } else {
                a.out[i] = total;
            }
        }
        if (a.comm_world > 1) {
            __threadfence_system();
            __syncthreads();
            if (tid < a.comm_world) {
                unsigned int* flag = reinterpret_cast<unsigned int*>(a.comm_peers[tid] + PINN_COMM_FLAGS_OFF) +
                                     par * PINN_COMM_MAX_RANKS + a.comm_rank;
                st_release_sys(flag, epoch);
                const unsigned int* mine_f = reinterpret_cast<const unsigned int*>(a.comm_peers[a.comm_rank] + PINN_COMM_FLAGS_OFF) +
                                             par * PINN_COMM_MAX_RANKS + tid;
                // a peer that never arrives (crashed rank) must not hang the GPU: give up after the time limit, and make
                // that sticky (word 1 of the local buffer) so that the remaining steps fail fast
                volatile unsigned int* aborted = reinterpret_cast<volatile unsigned int*>(a.comm_peers[a.comm_rank]) + 1;
                const long long t0 = clock64();
                while (ld_acquire_sys(mine_f) != epoch) {
                    if (*aborted || clock64() - t0 > a.comm_timeout) { *aborted = 1u; s_timeout = 1; break; }
                }
            }
            __syncthreads();
            const float* slots = slot_base + (size_t)par * a.comm_world * slot_f;
            for (int i = tid; i < n_out_floats; i += blockDim.x) {
                float s = 0.0f;
                for (int r = 0; r < a.comm_world; ++r) s += __ldcv(slots + (size_t)r * slot_f + i);
                a.out[i] = s_timeout ? __int_as_float(0x7fc00000) : s;      // poison instead of hanging
            }
        }
        if (tid == 0) *a.ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------
// The fit-step kernel.
// ---------------------------------------------------------------------------------------------------
template <int NF, int NS, bool GMEM, int MAXT, int JF, bool GEN>
__global__ void __launch_bounds__(MAXT, 1) step_kernel(const __grid_constant__ DevPlan P, const StepArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    const int n_out_floats = P.n_params + 4;
    const SmemLayout SL = smem_layout(P.weights_floats, n_out_floats, a.n_wacc,
                                      GMEM ? P.n_params : max(P.n_params, a.rows_total * RS * nwarps));
    stage_weights(smem, SL, P, a.params);
    const float* sw = smem + SL.weights_f;
    float* wacc_all = smem + SL.wacc_f;
    for (int i = tid; i < n_out_floats * a.n_wacc; i += blockDim.x) wacc_all[i] = 0.0f;
    __syncthreads();

    GradSink sink;
    sink.atomic = (a.n_wacc == 1 && nwarps > 1);
    sink.wacc = wacc_all + (a.n_wacc == 1 ? 0 : warp * n_out_floats);
    sink.dump = P.n_params + 2;                    // spare float behind the loss slot

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
        float r = point_step<NF, NS, JF, GEN>(P, sw, a.params, st, RS, valid, a.inv_n, sink, part);
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
    finish_grid(a, n_out_floats);
}


typedef void (*StepKernelFn)(const DevPlan, const StepArgs);

struct Variant {
    int nf, ns;
    StepKernelFn smem_fn, gmem_fn;           // plain problems (no residual wiring / oblique directions / IC variables)
    StepKernelFn smem_gen_fn, gmem_gen_fn;   // everything
    int maxt;
};

template <int NF, int NS>
struct VariantCfg {
    static constexpr int C = 1 + NF + NS;
    static constexpr int MAXT = (C <= 3) ? 512 : 256;
    static constexpr int JF = 16;
};

// `gen` selects which half of the function table this translation unit instantiates (the other half is
// filled by its sibling unit), so that the two halves compile in parallel.
template <int NF, int NS, bool GEN>
static Variant make_variant() {
    using Cfg = VariantCfg<NF, NS>;
    Variant v;
    v.nf = NF; v.ns = NS;
    v.smem_fn = v.gmem_fn = v.smem_gen_fn = v.gmem_gen_fn = nullptr;
    if (GEN) {
        v.smem_gen_fn = step_kernel<NF, NS, false, Cfg::MAXT, Cfg::JF, GEN>;
        v.gmem_gen_fn = step_kernel<NF, NS, true, Cfg::MAXT, Cfg::JF, GEN>;
    } else {
        v.smem_fn = step_kernel<NF, NS, false, Cfg::MAXT, Cfg::JF, GEN>;
        v.gmem_fn = step_kernel<NF, NS, true, Cfg::MAXT, Cfg::JF, GEN>;
    }
    v.maxt = Cfg::MAXT;
    return v;
}

}  // namespace pinn

// one lookup function per NF, each defined in its own translation unit
const pinn::Variant* pinn_variants_nf0(int ns);
const pinn::Variant* pinn_variants_nf1(int ns);
const pinn::Variant* pinn_variants_nf2(int ns);
const pinn::Variant* pinn_variants_nf3(int ns);
const pinn::Variant* pinn_variants_nf4(int ns);
const pinn::Variant* pinn_variants_gen_nf0(int ns);
const pinn::Variant* pinn_variants_gen_nf1(int ns);
const pinn::Variant* pinn_variants_gen_nf2(int ns);
const pinn::Variant* pinn_variants_gen_nf3(int ns);
const pinn::Variant* pinn_variants_gen_nf4(int ns);
