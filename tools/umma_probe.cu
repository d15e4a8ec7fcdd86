// umma_probe.cu — one-off hardware probe for the tcgen05 (UMMA) conventions the wide-net kernel relies on.
// There is no GPU in the build container, so every assumption about shared-memory matrix descriptors, the
// instruction descriptor, TMEM accumulator layouts and TMEM-sourced A operands is checked here against a CPU
// product before the real kernel is written.  Usage: umma_probe <test-id>   (each test in its own process, so a
// faulting variant cannot take the others down).  Build: nvcc -gencode arch=compute_100a,code=sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

enum { SRC_K_SW128 = 0, SRC_TMEM = 1, SRC_MN_SW128_32B = 2, SRC_MN_INTER = 3, SRC_K_INTER = 4, SRC_MN_SW128_16B = 5 };

struct Cfg {
    int M, N, K;
    int a_src, b_src;
    int d_lane_off;       // TMEM lane offset of the accumulator (0 or 16)
    int terms;            // 1: plain tf32; 3: 3xTF32 (hi/lo split on the device)
    int swap_lbo_sbo;     // MN-major variants: swap the two strides
    int reps;             // > 0: throughput test, issue `reps` x the k-loop
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                         // descriptor version 1 (Blackwell)
    d |= (uint64_t)(layout_type & 7) << 61;
    return d;
}

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// byte offset of logical element (row, k) of an operand stored in layout `src` (rows = M or N)
__device__ __forceinline__ uint32_t elem_off(int src, int rows, int K, int r, int k) {
    switch (src) {
        case SRC_K_SW128: {      // [k/32][r/8][r%8][128 B], 16-byte chunks XOR (r%8)
            return (uint32_t)((k / 32) * (rows / 8) * 1024 + (r / 8) * 1024 + (r % 8) * 128 + ((((k % 32) / 4) ^ (r % 8)) * 16) + (k % 4) * 4);
        }
        case SRC_K_INTER: {      // core matrix 8 rows x 16 B; [k/4][r/8][r%8][16 B]
            return (uint32_t)((k / 4) * (rows / 8) * 128 + (r / 8) * 128 + (r % 8) * 16 + (k % 4) * 4);
        }
        case SRC_MN_SW128_32B: { // [r/32][k][128 B], 32-byte chunks XOR (k%4)
            return (uint32_t)((r / 32) * K * 128 + k * 128 + ((((r % 32) / 8) ^ (k % 4)) * 32) + (r % 8) * 4);
        }
        case SRC_MN_SW128_16B: { // [r/32][k][128 B], 16-byte chunks XOR (k%8)
            return (uint32_t)((r / 32) * K * 128 + k * 128 + ((((r % 32) / 4) ^ (k % 8)) * 16) + (r % 4) * 4);
        }
        case SRC_MN_INTER: {     // core matrix 8 k-rows x 16 B (4 r's); [k/8][r/4][k%8][16 B]
            return (uint32_t)((k / 8) * (rows / 4) * 128 + (r / 4) * 128 + (k % 8) * 16 + (r % 4) * 4);
        }
    }
    return 0;
}

// descriptor of the K=8 slice starting at k0
__device__ __forceinline__ uint64_t operand_desc(int src, uint32_t base, int rows, int K, int k0, int swap) {
    switch (src) {
        case SRC_K_SW128:
            return make_desc(base + (k0 / 32) * (rows / 8) * 1024 + (k0 % 32) * 4, 16, 1024, 2);
        case SRC_K_INTER: {
            uint32_t lbo = (rows / 8) * 128, sbo = 128;
            return make_desc(base + (k0 / 4) * lbo, lbo, sbo, 0);
        }
        case SRC_MN_SW128_32B: {
            uint32_t lbo = K * 128, sbo = 512;
            if (swap) { uint32_t t = lbo; lbo = sbo; sbo = t; }
            return make_desc(base + k0 * 128, lbo, sbo, 1);
        }
        case SRC_MN_SW128_16B: {
            uint32_t lbo = K * 128, sbo = 1024;
            if (swap) { uint32_t t = lbo; lbo = sbo; sbo = t; }
            return make_desc(base + k0 * 128, lbo, sbo, 2);
        }
        case SRC_MN_INTER: {
            uint32_t sbo = 128, lbo = (rows / 4) * 128;
            if (swap) { uint32_t t = lbo; lbo = sbo; sbo = t; }
            return make_desc(base + (k0 / 8) * (rows / 4) * 128, lbo, sbo, 0);
        }
    }
    return 0;
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Cfg c, const float* A, const float* B, float* D, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int a_bytes = 128 * c.K * 4, b_bytes = ((c.N * c.K * 4) + 1023) & ~1023;
    uint8_t* sA[2] = {smem, smem + a_bytes};
    uint8_t* sB[2] = {smem + 2 * a_bytes, smem + 2 * a_bytes + b_bytes};

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the operand area, then fill (hi part in buffer 0, lo part in buffer 1)
    for (int i = tid; i < (2 * a_bytes + 2 * b_bytes) / 4; i += 128) ((float*)smem)[i] = 0.0f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t a_tmem_col = 256;                 // TMEM-sourced A lives in columns 256.. (hi), 256+K.. (lo)

    if (c.a_src != SRC_TMEM) {
        for (int i = tid; i < c.M * c.K; i += 128) {
            int r = i / c.K, k = i % c.K;
            float x = A[i], hi = (c.terms == 3) ? tf32_rn(x) : x;
            uint32_t off = elem_off(c.a_src, c.M, c.K, r, k);
            *(float*)(sA[0] + off) = hi;
            *(float*)(sA[1] + off) = x - hi;
        }
    } else {
        // thread = TMEM lane = row; columns = k
        for (int k0 = 0; k0 < c.K; k0 += 8) {
            uint32_t h[8], l[8];
            for (int j = 0; j < 8; ++j) {
                float x = (tid < c.M) ? A[tid * c.K + k0 + j] : 0.0f;
                float hi = (c.terms == 3) ? tf32_rn(x) : x;
                h[j] = __float_as_uint(hi); l[j] = __float_as_uint(x - hi);
            }
            uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + a_tmem_col + k0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "r"(ta), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
            ta += c.K;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "r"(ta), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < c.N * c.K; i += 128) {
        int r = i / c.K, k = i % c.K;
        float x = B[i], hi = (c.terms == 3) ? tf32_rn(x) : x;
        uint32_t off = elem_off(c.b_src, c.N, c.K, r, k);
        *(float*)(sB[0] + off) = hi;
        *(float*)(sB[1] + off) = x - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) |
                           ((c.a_src == SRC_MN_SW128_32B || c.a_src == SRC_MN_INTER || c.a_src == SRC_MN_SW128_16B) ? (1u << 15) : 0u) |
                           ((c.b_src == SRC_MN_SW128_32B || c.b_src == SRC_MN_INTER || c.b_src == SRC_MN_SW128_16B) ? (1u << 16) : 0u) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint32_t d_tmem = tmem + ((uint32_t)c.d_lane_off << 16);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        const int reps = c.reps > 0 ? c.reps : 1;
        uint32_t acc = 0;
        for (int rep = 0; rep < reps; ++rep) {
            // small terms first: lo*hi, hi*lo, then hi*hi
            for (int term = (c.terms == 3 ? 0 : 2); term < 3; ++term) {
                const int ai = (term == 0) ? 1 : 0, bi = (term == 1) ? 1 : 0;
                for (int k0 = 0; k0 < c.K; k0 += 8) {
                    uint64_t db = operand_desc(c.b_src, smem_u32(sB[bi]), c.N, c.K, k0, c.swap_lbo_sbo);
                    if (c.a_src == SRC_TMEM) {
                        mma_ts(d_tmem, tmem + a_tmem_col + ai * c.K + k0, db, idesc, acc);
                    } else {
                        uint64_t da = operand_desc(c.a_src, smem_u32(sA[ai]), c.M, c.K, k0, c.swap_lbo_sbo);
                        mma_ss(d_tmem, da, db, idesc, acc);
                    }
                    acc = 1;
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&s_bar)) : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0u) : "memory");
        }
    }
    if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // dump all 128 lanes x N columns
    for (int c0 = 0; c0 < c.N; c0 += 8) {
        uint32_t v[8];
        uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(ta) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[tid * c.N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

static const char* src_name(int s) {
    switch (s) { case 0: return "smem K-major SW128"; case 1: return "TMEM"; case 2: return "smem MN-major SW128_32B";
                 case 3: return "smem MN-major no-swizzle"; case 4: return "smem K-major no-swizzle"; case 5: return "smem MN-major SW128(16B)"; }
    return "?";
}

static int run(const char* title, Cfg c, bool ints) {
    printf("=== %s: M=%d N=%d K=%d A=%s B=%s d_lane_off=%d terms=%d swap=%d reps=%d\n", title, c.M, c.N, c.K, src_name(c.a_src),
           src_name(c.b_src), c.d_lane_off, c.terms, c.swap_lbo_sbo, c.reps);
    std::vector<float> A(c.M * c.K), B(c.N * c.K), D(128 * c.N, -777.0f);
    srand(1234);
    for (int r = 0; r < c.M; ++r) for (int k = 0; k < c.K; ++k)
        A[r * c.K + k] = ints ? (float)(((r * 7 + k * 3) % 11) - 5) : (float)rand() / RAND_MAX * 2.0f - 1.0f;
    for (int n = 0; n < c.N; ++n) for (int k = 0; k < c.K; ++k)
        B[n * c.K + k] = ints ? (float)(((n * 5 + k) % 7) - 3) : (float)rand() / RAND_MAX * 2.0f - 1.0f;
    float *dA, *dB, *dD; long long* dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
    int smem = 2 * 128 * c.K * 4 + 2 * (((c.N * c.K * 4) + 1023) & ~1023) + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(c, dA, dB, dD, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("RESULT %s: KERNEL FAULT %s\n", title, cudaGetErrorString(e)); return 1; }
    long long cyc = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
    const int reps = c.reps > 0 ? c.reps : 1;
    const int n_mma = reps * (c.terms == 3 ? 3 : 1) * (c.K / 8);
    printf("cycles %lld for %d MMAs -> %.1f cycles/MMA\n", cyc, n_mma, (double)cyc / n_mma);
    if (c.reps > 0) { printf("RESULT %s: TIMING ONLY\n", title); return 0; }
    // expected product in double
    std::vector<double> E(c.M * c.N);
    double emax = 0;
    for (int r = 0; r < c.M; ++r) for (int n = 0; n < c.N; ++n) {
        double s = 0; for (int k = 0; k < c.K; ++k) s += (double)A[r * c.K + k] * (double)B[n * c.K + k];
        E[r * c.N + n] = s; emax = fmax(emax, fabs(s));
    }
    // which row does every lane hold?
    int lane_row[128]; int matched = 0; double worst = 0;
    for (int lane = 0; lane < 128; ++lane) {
        lane_row[lane] = -1; double best = 1e30; int br = -1;
        for (int r = 0; r < c.M; ++r) {
            double err = 0; for (int n = 0; n < c.N; ++n) err = fmax(err, fabs((double)D[lane * c.N + n] - E[r * c.N + n]));
            if (err < best) { best = err; br = r; }
        }
        if (best <= (ints ? 1e-3 : (c.terms == 3 ? 2e-5 : 2e-2)) * fmax(1.0, emax)) { lane_row[lane] = br; ++matched; worst = fmax(worst, best / fmax(1.0, emax)); }
    }
    printf("lanes holding a correct row: %d of 128 (rows M=%d); worst rel err %.3e (emax %.3g)\n", matched, c.M, worst, emax);
    printf("lane->row:");
    for (int lane = 0; lane < 128; ++lane) { if (lane % 32 == 0) printf("\n  "); printf("%4d", lane_row[lane]); }
    printf("\n");
    if (matched < c.M) {
        printf("first lanes, first 8 columns (got | expected row=lane):\n");
        for (int lane = 0; lane < 4; ++lane) {
            for (int n = 0; n < 8; ++n) printf(" %9.3f", D[lane * c.N + n]);
            printf(" |");
            for (int n = 0; n < 8 && lane < c.M; ++n) printf(" %9.3f", E[lane * c.N + n]);
            printf("\n");
        }
    }
    bool ok = matched >= c.M;
    printf("RESULT %s: %s\n", title, ok ? "PASS" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    int t = argc > 1 ? atoi(argv[1]) : 0;
    switch (t) {
        case 0:  return run("ss_m128_k64", Cfg{128, 64, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 0}, true);
        case 1:  return run("ss_m128_3xtf32", Cfg{128, 64, 64, SRC_K_SW128, SRC_K_SW128, 0, 3, 0, 0}, false);
        case 2:  return run("ss_m128_1xtf32_random", Cfg{128, 64, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 0}, false);
        case 3:  return run("ss_m64_layout", Cfg{64, 64, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 0}, true);
        case 4:  return run("ss_m64_lane16", Cfg{64, 64, 64, SRC_K_SW128, SRC_K_SW128, 16, 1, 0, 0}, true);
        case 5:  return run("ts_m128", Cfg{128, 64, 64, SRC_TMEM, SRC_K_SW128, 0, 1, 0, 0}, true);
        case 6:  return run("ts_m128_3xtf32", Cfg{128, 64, 64, SRC_TMEM, SRC_K_SW128, 0, 3, 0, 0}, false);
        case 7:  return run("mnA_sw128_32b", Cfg{64, 64, 128, SRC_MN_SW128_32B, SRC_K_SW128, 0, 1, 0, 0}, true);
        case 8:  return run("mnA_sw128_32b_swap", Cfg{64, 64, 128, SRC_MN_SW128_32B, SRC_K_SW128, 0, 1, 1, 0}, true);
        case 9:  return run("mnA_inter", Cfg{64, 64, 128, SRC_MN_INTER, SRC_K_SW128, 0, 1, 0, 0}, true);
        case 10: return run("mnA_inter_swap", Cfg{64, 64, 128, SRC_MN_INTER, SRC_K_SW128, 0, 1, 1, 0}, true);
        case 11: return run("mnAB_sw128_32b", Cfg{64, 64, 128, SRC_MN_SW128_32B, SRC_MN_SW128_32B, 0, 1, 0, 0}, true);
        case 12: return run("k_inter", Cfg{128, 64, 64, SRC_K_INTER, SRC_K_INTER, 0, 1, 0, 0}, true);
        case 13: return run("mnA_sw128_16b", Cfg{64, 64, 128, SRC_MN_SW128_16B, SRC_K_SW128, 0, 1, 0, 0}, true);
        case 14: return run("mnA_sw128_16b_swap", Cfg{64, 64, 128, SRC_MN_SW128_16B, SRC_K_SW128, 0, 1, 1, 0}, true);
        case 15: return run("ss_m128_n256", Cfg{128, 256, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 0}, true);
        case 16: return run("ss_m128_n144", Cfg{128, 144, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 0}, true);
        // throughput
        case 20: return run("time_ss_m128_n64", Cfg{128, 64, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 200}, true);
        case 21: return run("time_ss_m64_n64", Cfg{64, 64, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 200}, true);
        case 22: return run("time_ts_m128_n64", Cfg{128, 64, 64, SRC_TMEM, SRC_K_SW128, 0, 1, 0, 200}, true);
        case 23: return run("time_ss_m128_n256", Cfg{128, 256, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 200}, true);
        case 24: return run("time_ss_m128_n128", Cfg{128, 128, 64, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 200}, true);
        case 25: return run("time_ss_m64_n128_k128", Cfg{64, 128, 128, SRC_K_SW128, SRC_K_SW128, 0, 1, 0, 100}, true);
    }
    printf("unknown test %d\n", t);
    return 3;
}
