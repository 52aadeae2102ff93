// mma_probe.cu -- bring-up tool (NOT part of the product library): how many cycles does one tcgen05.mma kind::f16 take on
// sm_100a as a function of N, accumulator switching, A-operand row shift (the implicit-GEMM tap offset of net_tc.cu) and the
// shared-memory layout (K-major no-swizzle "interleaved" vs SWIZZLE_128B)?  One CTA, one issuing thread, clock64 around
// [issue n MMAs -> tcgen05.commit -> mbarrier wait].  Operand values are irrelevant (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -I lightzero_b200/csrc tests/csrc/mma_probe.cu -o /tmp/libmmaprobe.so
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "tc_ptx.cuh"

using namespace lz;

struct ProbeCfg {
    int N;              // 32 / 64 / 128 / 256
    int n_mma;          // MMAs per measurement
    int n_acc;          // accumulators used round-robin ...
    int switch_every;   // ... switching after this many MMAs
    int a_shift_rows;   // start-address offset of A in rows (16 B each without swizzle, 128 B with SWIZZLE_128B)
    int a_lbo16;        // no-swizzle: k-group stride of A in 16-byte units (128 = 128-row tile, 400 = net_tc.cu's activation buffer)
    int swz_a, swz_b;   // 1: SWIZZLE_128B K-major (rows of 128 B, 8-row groups of 1024 B)
    int n_ksteps;       // distinct K-steps cycled through (operand addresses change every MMA like in the real kernels)
    int mix_n2;         // > 0: groups of `mix_group` MMAs alternate between N and this second N (net_tc.cu's [B_hi | B_lo] fold: N = 128 then N = 64)
    int mix_group;
    int mix_d2_off;     // TMEM column offset of the second shape's accumulator relative to the first (0: same columns)
    int warp_issue;     // 1: the whole warp runs the issue loop and the MMA is predicated on elect.sync (uniform control flow,
                        //    the CUTLASS pattern); 0: `if (lane == 0)` divergent branch (what net_tc.cu / conv_tc.cu / ez.cu do)
};

__device__ __forceinline__ uint64_t desc_swz128(uint32_t saddr)
{
    // K-major SWIZZLE_128B: LBO unused (1), SBO = 1024 B, layout type 2 at bits [61,64), base offset at [49,52) when the
    // start address is not 1024-byte aligned
    const uint64_t base_off = (uint64_t)((saddr >> 7) & 7u);
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (base_off << 49) | (2ull << 61);
}

__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    lz::umma_f16_elect(d_tmem, adesc, bdesc, idesc, 1u);
}

extern "C" __global__ void __launch_bounds__(128, 1) k_mma_probe(const ProbeCfg *cfgs, int ncfg, unsigned long long *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ uint64_t t_ad[512], t_bd[512];      // descriptors precomputed per configuration: the timed loop only issues
    __shared__ uint32_t t_d[512], t_id[512];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (warp == 0 && (cfgs[0].warp_issue || tid == 0)) {
        const bool wi = cfgs[0].warp_issue != 0;
        const uint32_t a_s = smem_u32(smem) + 16 * 1024, b_s = smem_u32(smem) + 120 * 1024;     // room for negative shifts
        uint32_t parity = 0;
        for (int c = 0; c < ncfg; ++c) {
            const ProbeCfg g = cfgs[c];
            const uint32_t idesc = make_idesc_f16(128, g.N);
            for (int i = (wi ? (tid & 31) : 0); i < g.n_mma; i += (wi ? 32 : 1)) {
                const int ks = i % g.n_ksteps, acc = (i / g.switch_every) % g.n_acc;
                if (g.swz_a) t_ad[i] = desc_swz128(a_s + (uint32_t)g.a_shift_rows * 128u + (uint32_t)ks * 32u);
                else t_ad[i] = make_desc(a_s + (uint32_t)g.a_shift_rows * 16u + (uint32_t)ks * 2u * (uint32_t)g.a_lbo16 * 16u, g.a_lbo16, 8);
                if (g.swz_b) t_bd[i] = desc_swz128(b_s + (uint32_t)ks * 32u);
                else t_bd[i] = make_desc(b_s + (uint32_t)ks * 2u * (uint32_t)g.N * 16u, g.N, 8);
                t_d[i] = tmem + acc * g.N;
                t_id[i] = idesc;
                if (g.mix_n2 > 0 && ((i / g.mix_group) & 1)) {
                    t_id[i] = make_idesc_f16(128, g.mix_n2);
                    t_d[i] += g.mix_d2_off;
                }
            }
            if (wi) __syncwarp();
            for (int rep = 0; rep < 3; ++rep) {      // the last repetition is reported
                const long long t0 = clock64();
                if (wi) {
#pragma unroll 4
                    for (int i = 0; i < g.n_mma; ++i) umma_f16_elect(t_d[i], t_ad[i], t_bd[i], t_id[i]);
                    umma_commit_elect(&bar);
                } else {
#pragma unroll 4
                    for (int i = 0; i < g.n_mma; ++i) umma_f16(t_d[i], t_ad[i], t_bd[i], t_id[i], 1);
                    umma_commit(&bar);
                }
                mbar_wait(&bar, parity);
                parity ^= 1;
                const long long t1 = clock64();
                if (tid == 0) out[c] = (unsigned long long)(t1 - t0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

extern "C" int mma_probe_run(const ProbeCfg *h_cfgs, int ncfg, unsigned long long *h_out)
{
    // one launch per configuration, results copied out immediately: a faulting descriptor only loses the configurations after it
    ProbeCfg *d_cfg = nullptr;
    unsigned long long *d_out = nullptr;
    if (cudaMalloc(&d_cfg, sizeof(ProbeCfg)) != cudaSuccess || cudaMalloc(&d_out, 8) != cudaSuccess) return -1;
    const int smem = 200 * 1024;
    if (cudaFuncSetAttribute(k_mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
    int done = 0;
    for (int c = 0; c < ncfg; ++c, ++done) {
        cudaMemcpy(d_cfg, h_cfgs + c, sizeof(ProbeCfg), cudaMemcpyHostToDevice);
        cudaMemset(d_out, 0, 8);
        k_mma_probe<<<1, 128, smem>>>(d_cfg, 1, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { fprintf(stderr, "mma_probe: configuration %d: %s\n", c, cudaGetErrorString(e)); break; }
        cudaMemcpy(h_out + c, d_out, 8, cudaMemcpyDeviceToHost);
    }
    return done;
}
