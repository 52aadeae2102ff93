// tc_ptx.cuh -- thin inline-PTX wrappers for the Blackwell (sm_100a) primitives used by the tcgen05 kernels:
// mbarrier, cp.async.bulk (TMA bulk copy), proxy / tcgen05 fences, TMEM alloc / ld / st, tcgen05.mma
// (kind::f16, cta_group::1) and tcgen05.commit, plus the UMMA shared-memory and instruction descriptors.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace lz {

// Programmatic dependent launch: wait for the prerequisite grid(s) / let dependents start their prologue early
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 28); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    printf("lz net_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
    asm volatile("trap;\n");
}
// Whole-warp wait: only lane 0 polls (with back-off) so that idle warps do not compete with the tensor
// core's operand fetch for shared-memory bandwidth; the other lanes park at the warp barrier.
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity)
{
    if ((threadIdx.x & 31) == 0) {
        const uint32_t a = smem_u32(bar);
        uint32_t ok = 0;
        for (uint32_t it = 0; it < (1u << 26) && !ok; ++it) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(a), "r"(parity) : "memory");
#ifndef LZ_POLL_NS
#define LZ_POLL_NS 64
#endif
            if (!ok) __nanosleep(LZ_POLL_NS);
        }
        if (!ok) {
            printf("lz net_tc: mbarrier timeout (block %d warp %d)\n", blockIdx.x, threadIdx.x >> 5);
            asm volatile("trap;\n");
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 inputs, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// The same, to be executed by ALL lanes of the issuing warp in uniform control flow: one elected lane issues.  Measured
// (profiles/r01e_mma_probe.md): 48.6 cycles per N = 64 MMA (the shared-memory operand floor) against 60-78 when the
// instruction sits in an `if (lane == 0)` branch, where ptxas wraps it in an ELECT / BRA.U.ANY loop.
__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred pe;\n\t.reg .pred pa;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
                 "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar)
{
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// one elected lane of a converged warp (true in exactly one lane)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// whole-warp wait for a warp that must stay converged (the MMA issuer): one lane polls, the others park at the warp barrier
__device__ __forceinline__ void mbar_wait_converged(uint64_t *bar, uint32_t parity)
{
#ifdef LZ_WAIT_ONE_LANE
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
#else
    // every lane polls (same address: one shared-memory access per try): no divergent branch at all in front of the elect.sync
    // issue loop -- a lane-0-only poll followed by __syncwarp() left the warp split and every elect.sync re-converging (measured:
    // ~330 cycles per MMA instead of 49-65)
    mbar_wait(bar, parity);
    __syncwarp();
#endif
}
// tcgen05.ld without the wait (several loads in flight), and the wait that also pins the destination registers so the
// compiler cannot schedule their uses above it
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tmem_pin(uint32_t (&r)[N])
{
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(r[i]));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
                 ::"r"(taddr),
                   "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
                   "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
                   "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
                   "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
                   "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), leading byte offset >> 4 in [16,30) (between the two 16-byte K chunks of
// one MMA), stride byte offset >> 4 in [32,46) (between 8-row core matrices), version = 1 in [46,48).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = F16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// split 8 floats into fp16 hi / lo and store them as two 16-byte vectors
__device__ __forceinline__ void store_split8(unsigned char *hi_ptr, unsigned char *lo_ptr, const float *v)
{
    __half2 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a = v[2 * i], b = v[2 * i + 1];
        a = fminf(fmaxf(a, -65504.0f), 65504.0f);
        b = fminf(fmaxf(b, -65504.0f), 65504.0f);
        __half ha = __float2half_rn(a), hb = __float2half_rn(b);
        h[i] = __halves2half2(ha, hb);
        l[i] = __halves2half2(__float2half_rn(a - __half2float(ha)), __float2half_rn(b - __half2float(hb)));
    }
    *reinterpret_cast<uint4 *>(hi_ptr) = *reinterpret_cast<uint4 *>(h);
    *reinterpret_cast<uint4 *>(lo_ptr) = *reinterpret_cast<uint4 *>(l);
}


// the same for values known to be >= 0 (post-ReLU activations): only the upper clamp, packed conversions
__device__ __forceinline__ void store_split8_pos(unsigned char *hi_ptr, unsigned char *lo_ptr, const float *v)
{
    __half2 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = fminf(v[2 * i], 65504.0f), b = fminf(v[2 * i + 1], 65504.0f);
        h[i] = __floats2half2_rn(a, b);
        const float2 back = __half22float2(h[i]);
        l[i] = __floats2half2_rn(a - back.x, b - back.y);
    }
    *reinterpret_cast<uint4 *>(hi_ptr) = *reinterpret_cast<uint4 *>(h);
    *reinterpret_cast<uint4 *>(lo_ptr) = *reinterpret_cast<uint4 *>(l);
}

}  // namespace lz
