// net_tc.cu -- tcgen05 (5th-gen tensor core) path of the MuZero latent-grid networks for sm_100a.
//
// One CTA runs the WHOLE recurrent_inference (or the latent-grid tail of initial_inference) for up to 7 roots -- and, in persistent
// mode, the whole num_simulations loop of their search (tree back-up + descent by one warp per tree, tree_persist.cuh): five 3x3
// convolutions, three 1x1 head convolutions and the heads' fully connected layers as tcgen05.mma with fp32 accumulators in TMEM;
// BatchNorm / residual / ReLU epilogues and the softmax expectation + inverse scalar transform straight out of TMEM.  Activations never
// leave the SM.  DESIGN.md 4.3 / 4.3c describe the design and what bounds it (the L1 / shared-memory data pipe).
//
// fp32 accuracy on fp16 tensor cores ("3xFP16"): every fp32 operand v is split v = hi + lo with hi = fp16(v), lo = fp16(v - hi) (22
// significant bits); D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi with fp32 accumulation drops only the 2^-22 lo*lo term.  The tap block of the
// weights stores B_hi and B_lo as 128 consecutive operand rows, so A_hi x [B_hi | B_lo] is ONE N = 128 MMA and A_lo x B_hi a second one of
// N = 64; the two 64-column halves of a tile's accumulator are added at read-out.  Weights are pre-split on the host and pre-scaled by a
// power of two (exact; folded back into the BatchNorm scale) so their lo parts stay in fp16's normal range.  Mode 2 ("fast") issues only
// the hi*hi pass.
//
// Implicit GEMM without im2col: activations live in shared memory as [k-group of 8 channels][row][8 halves] (the UMMA K-major no-swizzle
// canonical layout with SBO = 128 B, so row r of the operand is at start + 16*r bytes).  Rows are the pixels of a 7-wide padded grid (49
// rows per root, column 6 and row 6 zero), so the input of output row m for tap (dy,dx) is row m + 7*dy + dx: each of the 9 taps is the
// SAME buffer addressed through a descriptor whose start address is shifted by (7*dy+dx)*16 bytes.  The zero pad rows double as the conv
// padding between rows and between consecutive roots.  M tiles of 128 rows cut anywhere; all tiles of a layer (of a root group, see the
// kernel) accumulate in TMEM before the epilogue rewrites the buffer IN PLACE; the ResBlock skip tensors are parked in an L2-resident
// scratch (thread-private rows, .cg accesses) -- TMEM is full: 3 tiles x 128 accumulator columns + the reward hook.
//
// Warp roles (320 threads): warps 0-7 = epilogue / loads / heads / trees (warp w owns TMEM lanes 32*(w%4).. and the 32-column half w/4 of
// every accumulator), warp 8 lane 0 = weight producer (cp.async.bulk global->shared ring, mbarrier complete_tx), warp 9 = MMA issuer (the
// whole warp runs the issue loops in uniform control flow, elect.sync picks the lane) + TMEM alloc/dealloc.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "model.cuh"
#include "net_tc.cuh"
#include "tc_ptx.cuh"
#include "tree.cuh"
#include "tree_persist.cuh"

namespace lz {


// ---------------------------------------------------------------------------------------------- geometry
constexpr int kEpiWarps = 8, kEpiThreads = kEpiWarps * 32;   // two warps per TMEM lane quarter, one 32-column half each
constexpr int kTcThreads = kEpiThreads + 64;
constexpr int kPitch = 7, kRowsPerRoot = 49;     // padded 7x7 grid per root
constexpr int kMaxRoots = 7, kMaxTiles = 3;      // 343 rows -> 3 tiles of 128
constexpr int kMargin = 8;                       // |7*dy+dx| <= 8
constexpr int kRowsAlloc = kMargin + kMaxTiles * 128 + kMargin;   // 400
constexpr int kPlaneBytes = kRowsAlloc * 16;     // one k-group (8 fp16 channels) of all rows: 6400 B
constexpr int kPartBytes = 8 * kPlaneBytes;      // 64 channels: 51200 B
constexpr int kActBytes = 2 * kPartBytes;        // hi + lo: 102400 B
constexpr int kTapKgBytes = 128 * 16;            // one k-group (8 input channels) of a tap: 64 hi rows then 64 lo rows of 16 B
constexpr int kTapBytes = 8 * kTapKgBytes;       // one 3x3 tap, [kg 8][co: 64 hi | 64 lo][ci % 8]: 16384 B
constexpr int kStages = 5;                       // 16 KB ring stages: conv taps, then the heads' FC weight blocks
constexpr int kHeadWBytes = 3 * 2 * 16 * 64 * 2; // three 1x1 heads (hc <= 16), hi + lo: 12288 B
constexpr int kBnSmemBytes = kTcMaxLayers * 128 * 4;   // folded BatchNorm tables of the program's layers
constexpr int kSmemMain = kActBytes + kStages * kTapBytes + kHeadWBytes + 1024 + kBnSmemBytes;
constexpr int kHeadScratch = 72 * 8 * 16 * 2 + kEpiWarps * (16 + 3 * 32) * 4;   // kFrBytes (reward features parked from their hook to the heads' FC pass) + the parked trees (kTreeParkWords per warp)
// tree <-> network hand-off of the persistent search, per root slot of the CTA: leaf slot, action | value, reward, policy logits
// (the global copies are still written for the step-wise entry points; reading them back would cost an L2 round trip per use)
constexpr int kHoWords = 4 + 32;
constexpr int kHandoffBytes = 8 * kHoWords * 4;
constexpr int kSmemBytes = kSmemMain + kHeadScratch + kHandoffBytes;

// TMEM columns
constexpr int kColAcc = 0;        // 3 tiles x 128: [0,64) = A_hi*B_hi + A_lo*B_hi, [64,128) = A_hi*B_lo (one N = 128 MMA)
constexpr int kAccCols = 128;
constexpr int kColVp = 0;         // 3 tiles x 32  (value + policy 1x1: reuses the drained accumulator columns of tile 0)
constexpr int kColRew = 384;      // 3 tiles x 16  (reward 1x1)
constexpr int kTmemCols = 512;

struct TcBars {
    uint64_t full[kStages], empty[kStages];
    uint64_t acc_ready[2];  // MMA -> epilogue: this layer's accumulators of root group g are complete
    uint64_t act_ready[2];  // epilogue -> MMA: activations (and TMEM) of root group g are ready for the next layer
    uint64_t rew_ready;     // MMA -> heads: reward 1x1 accumulators complete (one phase per simulation, one arrival per root group)
    uint64_t vp_ready;      // MMA -> heads: value/policy 1x1 accumulators complete (likewise)
    uint64_t fcb_ready;     // heads -> MMA: the head features (FC1's B operand) are in shared memory
    uint64_t fc1_done;      // MMA -> heads: FC1 accumulators complete
    uint64_t fc2b_ready;    // heads -> MMA: the hidden activations (FC2's B operand) are in shared memory
    uint64_t fc2_done;      // MMA -> heads: FC2 accumulators (the heads' logits) complete
    uint32_t tmem_base;
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------- heads
// The fully connected parts of the heads (reward / value / policy: Linear(hc*36 -> hid) + BN + ReLU, Linear(hid -> K),
// muzero_model.py:465-502, common.py:1130-1187) run on the TENSOR CORES at the end of a simulation with the roles swapped: the
// weights are the M operand (streamed through the same 16 KB ring as the conv taps, fp16 hi/lo), the 7 roots are N columns:
//   FC1  D[(head, unit) 96 of 128 rows][(head, root) 32 columns] += W1^T[rows][K = 576 inputs] x F[K][columns]; only the diagonal
//        (head == head) blocks are read back (one MMA stream for all three heads, A-operand-bound);
//   FC2  per head and per 128-output tile: D[output k][root] = W2[k][K = 32 units] x H[units][root]: a thread of the read-out owns
//        output k = tile * 128 + lane for every root -- exactly the distribution of the canonical softmax order (net6.cuh), so the
//        softmax expectation runs straight out of TMEM, the 601 logits are never stored.
// Both products keep the fp32-accurate 3xFP16 scheme: A_hi x [B_hi | B_lo] and A_lo x [B_hi | B_lo] (the extra lo x lo term is
// harmless), the two column halves are added at read-out.
constexpr int kFbKgBytes = 64 * 16 + 16;                  // FC1 B operand: per k-group 32 hi rows + 32 lo rows of 16 B (+16 B: bank spread)
constexpr int kFbBytes = 72 * kFbKgBytes;                 // 576 inputs = 72 k-groups: 74,880 B (overlays the activation buffer)
constexpr int kFrBytes = 72 * 8 * 16 * 2;                 // reward features parked from their hook: [hi | lo][72 k-groups][8 roots][16 B] = 18,432 B
constexpr int kHbHeadBytes = 4 * 16 * 16;                 // FC2 B operand of one head: [4 k-groups][8 roots hi | 8 roots lo][16 B] = 1 KB
constexpr int kFc1Stages = 18;                            // 576 inputs / 32 per stage (2 k-steps x (A_hi + A_lo))
constexpr int kFc1StageBytes = 2 * 2 * 2 * 96 * 16;       // 12,288 B: only the 96 real rows (3 heads x 32 units) of the M = 128 operand are streamed
// TMEM columns of the FC phase (the conv accumulators are drained by then; the value/policy 1x1 results in columns 0-95 are
// consumed by head_scatter before FC1 starts)
constexpr int kColFc1 = 0;        // 64: [0,32) hi x hi + lo x hi, [32,64) hi x lo (+ lo x lo)
constexpr int kColFc2 = 64;       // 16 per (head, 128-output tile): [0,8) / [8,16) likewise

__device__ __forceinline__ void put_half(unsigned char *p, float v, float &rem)
{
    const __half h = __float2half_rn(fminf(v, 65504.0f));
    *reinterpret_cast<__half *>(p) = h;
    rem = v - __half2float(h);
}

// 1x1-conv accumulators (TMEM) -> BatchNorm + ReLU -> fp16 hi/lo features in FC1's B-operand layout, for the heads in hmask (bit 0
// reward -> its parking buffer fr, bit 1 value / bit 2 policy -> rows 8-15 / 16-23 of fb).  Feature k = c * 36 + p of root r sits at
// k-group k / 8, row (head * 8 + r), element k % 8.  No barrier inside.
// Row m of the CTA's padded pixel grid -> (root slot r in the CTA, pixel p); false for pad rows / absent roots.  Root group X
// (roots 0 .. Rx-1) starts at row 0, root group Y (roots Rx .. R-1) at row ybase (a tile boundary); without a split Rx = R.
struct RowMap { int Rx, R, ybase, nvalid; };
__device__ __forceinline__ bool row_decode(const RowMap &rm, int m, int &r, int &p)
{
    const bool in_y = m >= rm.ybase;
    const int mm = in_y ? m - rm.ybase : m;
    const int rr = mm / kRowsPerRoot, q = mm - rr * kRowsPerRoot, y = q / kPitch, x = q - y * kPitch;
    r = in_y ? rm.Rx + rr : rr;
    p = y * 6 + x;
    return (in_y ? (r < rm.R) : (rr < rm.Rx)) && (r < rm.nvalid) && (y < 6) && (x < 6);
}

__device__ __forceinline__ void head_scatter(const TcNet &net, int hmask, unsigned char *fr, unsigned char *fb, uint32_t tmem, int NT,
                                             const RowMap &rm)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q4 = warp & 3, half = warp >> 2, rowid = q4 * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    for (int t = 0; t < NT; ++t) {
        int r, p;
        const bool valid = row_decode(rm, t * 128 + rowid, r, p);
        float v[16];
        // warps 0-3 scatter reward + value features, warps 4-7 policy features (any warp may read its lane quarter).  The channel
        // loops are unrolled over the 16 accumulator columns (v[] must stay in registers) and predicated on the head's width.
        if (half == 0 && (hmask & 1)) {
            tmem_ld16(lane_base + kColRew + t * 16, v);
            if (valid) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (c >= net.hc[0]) break;
                    const int k = c * kP + p;
                    unsigned char *d = fr + ((k >> 3) * 8 + r) * 16 + (k & 7) * 2;
                    float rem;
                    put_half(d, fmaxf(fmaf(v[c], __ldg(net.head_bn + c), __ldg(net.head_bn + 16 + c)), 0.0f), rem);
                    *reinterpret_cast<__half *>(d + kFrBytes / 2) = __float2half_rn(rem);
                }
            }
        }
        if (half == 0 && (hmask & 2)) {
            tmem_ld16(lane_base + kColVp + t * 32, v);
            if (valid) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (c >= net.hc[1]) break;
                    const int k = c * kP + p;
                    unsigned char *d = fb + (k >> 3) * kFbKgBytes + (8 + r) * 16 + (k & 7) * 2;
                    float rem;
                    put_half(d, fmaxf(fmaf(v[c], __ldg(net.head_bn + 32 + c), __ldg(net.head_bn + 48 + c)), 0.0f), rem);
                    *reinterpret_cast<__half *>(d + 32 * 16) = __float2half_rn(rem);
                }
            }
        }
        if (half == 1 && (hmask & 4)) {
            tmem_ld16(lane_base + kColVp + t * 32 + 16, v);
            if (valid) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (c >= net.hc[2]) break;
                    const int k = c * kP + p;
                    unsigned char *d = fb + (k >> 3) * kFbKgBytes + (16 + r) * 16 + (k & 7) * 2;
                    float rem;
                    put_half(d, fmaxf(fmaf(v[c], __ldg(net.head_bn + 64 + c), __ldg(net.head_bn + 80 + c)), 0.0f), rem);
                    *reinterpret_cast<__half *>(d + 32 * 16) = __float2half_rn(rem);
                }
            }
        }
    }
    tc_fence_before();
}

// FC1 read-out (warps 0-2: rows 32 h + unit j of head h): BatchNorm + ReLU, then the hidden activations as FC2's B operand
// hb[head][k-group j / 8][root | 8 + root (lo)][j % 8].
__device__ __forceinline__ void heads_hidden(const TcNet &net, int hmask, unsigned char *hb, uint32_t tmem)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 3 || !((hmask >> warp) & 1)) return;
    const int h = warp, j = lane;
    const Head &H = h == 0 ? net.reward : (h == 1 ? net.value : net.policy);
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t a[8], b[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "r"(lane_base + kColFc1 + h * 8));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "r"(lane_base + kColFc1 + 32 + h * 8));
    tmem_ld_wait();
    tmem_pin(a); tmem_pin(b);
    const float inv = net.fc[h].fc1_inv;
    const float s2 = (j < H.hid) ? __ldg(H.s2 + j) : 0.0f, t2 = (j < H.hid) ? __ldg(H.t2 + j) : 0.0f;
    unsigned char *dst = hb + h * kHbHeadBytes + (j >> 3) * 256 + (j & 7) * 2;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const float pre = (__uint_as_float(a[r]) + __uint_as_float(b[r])) * inv;
        const float hid = (j < H.hid) ? fmaxf(fmaf(pre, s2, t2), 0.0f) : 0.0f;
        float rem;
        put_half(dst + r * 16, hid, rem);
        *reinterpret_cast<__half *>(dst + (8 + r) * 16) = __float2half_rn(rem);
    }
    tc_fence_before();
}

// FC2 read-out: thread (g = root group of 4, wg = TMEM lane quarter, lane) owns output k = tile * 128 + wg * 32 + lane of roots
// 4 g .. 4 g + 3: logits = D / scale + bias; raw logits to global when asked for; the categorical heads fold them into the
// canonical one-pass softmax expectation (net6.cuh) and the inverse scalar transform.  Same arithmetic per (head, root) as
// categorical_to_scalar; the two categorical heads are reduced TOGETHER (their 8 max / 16 sum butterflies interleaved, two CTA
// barriers instead of six: the read-out is a latency chain, not a throughput problem).  red: 2 x 2 x 4 x 4 x 3 floats of scratch.
__device__ __forceinline__ void heads_outputs(const TcNet &net, const TcIO &io, int hmask, uint32_t tmem, float *red, int nvalid, int root0,
                                              const float *b2_s, float *ho)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = warp >> 2, wg = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(wg * 32) << 16);
    float m[2][4], sm[2][4], ws[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 0; q < 4; ++q) { m[h][q] = -INFINITY; sm[h][q] = 0.0f; ws[h][q] = 0.0f; }
    int tile = 0, boff = 0;
    int hdone = 0;
#ifndef LZ_NO_JOINT_HEADS
    if ((hmask & 3) == 3 && net.reward.K == net.value.K && !io.reward_logits && !io.value_logits) {
        // the two categorical heads of the search (same support size, no raw logits wanted) in ONE loop: 8 independent softmax
        // recurrences per thread instead of 4 -- the read-out is a chain of dependent latencies, so this nearly halves it
        const int K = net.reward.K, nblk = (K + 127) >> 7;
        const float inv0 = net.fc[0].fc2_inv, inv1 = net.fc[1].fc2_inv;
        const float *b20 = b2_s ? b2_s : net.reward.b2, *b21 = b2_s ? b2_s + K : net.value.b2;
#pragma unroll 1
        for (int blk = 0; blk < nblk; ++blk) {
            const int k = blk * 128 + wg * 32 + lane;
            const float bias0 = (k < K) ? b20[k] : 0.0f, bias1 = (k < K) ? b21[k] : 0.0f;
            uint32_t a0[4], c0[4], a1[4], c1[4];
            const uint32_t col0 = lane_base + kColFc2 + blk * 16 + g * 4, col1 = col0 + nblk * 16;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a0[0]), "=r"(a0[1]), "=r"(a0[2]), "=r"(a0[3]) : "r"(col0));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(c0[0]), "=r"(c0[1]), "=r"(c0[2]), "=r"(c0[3]) : "r"(col0 + 8));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a1[0]), "=r"(a1[1]), "=r"(a1[2]), "=r"(a1[3]) : "r"(col1));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(c1[0]), "=r"(c1[1]), "=r"(c1[2]), "=r"(c1[3]) : "r"(col1 + 8));
            tmem_ld_wait();
            tmem_pin(a0); tmem_pin(c0); tmem_pin(a1); tmem_pin(c1);
            if (k < K) {
                const float sup = support_at(net.support_min, net.support_step, k);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (g * 4 + q >= nvalid) continue;
                    softmax_push(m[0][q], sm[0][q], ws[0][q], fmaf(__uint_as_float(a0[q]) + __uint_as_float(c0[q]), inv0, bias0), sup);
                    softmax_push(m[1][q], sm[1][q], ws[1][q], fmaf(__uint_as_float(a1[q]) + __uint_as_float(c1[q]), inv1, bias1), sup);
                }
            }
        }
        tile = 2 * nblk; boff = 2 * K; hdone = 3;
    }
#endif
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        if (!((hmask >> h) & 1) || ((hdone >> h) & 1)) continue;
        const Head &H = h == 0 ? net.reward : (h == 1 ? net.value : net.policy);
        const int K = H.K, nblk = (K + 127) >> 7;
        const float inv = net.fc[h].fc2_inv;
        float *glog = h == 0 ? io.reward_logits : (h == 1 ? io.value_logits : io.policy_logits);
        // FC2 bias: the copy staged in shared memory at kernel start (the evaluated heads back to back) or global
        const float *b2 = b2_s ? b2_s + boff : H.b2;
        boff += K;
#pragma unroll 1
        for (int blk = 0; blk < nblk; ++blk, ++tile) {      // rolled: the read-out is a long latency chain, its code must stay small
            const int k = blk * 128 + wg * 32 + lane;
            const float bias = (k < K) ? b2[k] : 0.0f;
            uint32_t a[4], b[4];
            const uint32_t col = lane_base + kColFc2 + tile * 16 + g * 4;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(col));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(col + 8));
            tmem_ld_wait();
            tmem_pin(a); tmem_pin(b);
            if (k < K) {
                const float sup = support_at(net.support_min, net.support_step, k);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = g * 4 + q;
                    if (r >= nvalid) continue;
                    const float o = fmaf(__uint_as_float(a[q]) + __uint_as_float(b[q]), inv, bias);
                    if (glog) glog[(size_t)(root0 + r) * K + k] = o;
                    if (h == 2 && ho && k < 32) ho[r * kHoWords + 4 + k] = o;
                    if (h < 2) softmax_push(m[h & 1][q], sm[h & 1][q], ws[h & 1][q], o, sup);   // running softmax statistics (net6.cuh: the canonical order)
                }
            }
        }
    }
    // block-wide combine per (head, root): max, then rescaled sums (4 warps per root group); both heads at once
    float mg[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 0; q < 4; ++q) mg[h][q] = m[h][q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q) mg[h][q] = fmaxf(mg[h][q], __shfl_xor_sync(0xffffffffu, mg[h][q], o));
    auto red_at = [&](int h, int w4, int q) { return red + (((h * 2 + g) * 4 + w4) * 4 + q) * 3; };
    if (lane == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q) red_at(h, wg, q)[0] = mg[h][q];
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float M = red_at(h, 0, q)[0];
#pragma unroll
            for (int w4 = 1; w4 < 4; ++w4) M = fmaxf(M, red_at(h, w4, q)[0]);
            const float sc = (m[h][q] == -INFINITY) ? 0.0f : expf(m[h][q] - M);
            sm[h][q] *= sc;
            ws[h][q] *= sc;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                sm[h][q] += __shfl_xor_sync(0xffffffffu, sm[h][q], o);
                ws[h][q] += __shfl_xor_sync(0xffffffffu, ws[h][q], o);
            }
    if (lane == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q) { red_at(h, wg, q)[1] = sm[h][q]; red_at(h, wg, q)[2] = ws[h][q]; }
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
    if (wg == 0 && lane < 8) {
        const int h = lane >> 2, q = lane & 3, r = g * 4 + q;
        if (r < nvalid && ((hmask >> h) & 1)) {
            float S = 0.0f, W = 0.0f;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) { S += red_at(h, w4, q)[1]; W += red_at(h, w4, q)[2]; }
            const float v = inverse_scalar_transform(W / S);
            float *dst = h == 0 ? io.reward : io.value;
            if (dst) dst[root0 + r] = v;
            if (ho) ho[r * kHoWords + (h == 0 ? 3 : 2)] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------- kernel
// 32 consecutive channels [32*half, 32*half+32) of pixel p of one root latent.  cl: the kernel-internal layout [c / 4][36][c % 4]
// (the pool slots this kernel writes in persistent mode, the skip scratch, the action-bias table): a thread's float4 j is at
// ((c0 / 4 + j) * 36 + p) * 4, so the lanes of a warp (consecutive pixels) touch consecutive 16-byte chunks -- coalesced 512-byte
// warp accesses (a plain channels-last row per lane costs 32 separate sectors per warp instruction).  Else NCHW [64][36]
// (every tensor that crosses the API).
__device__ __forceinline__ void load_row32(const float *root, bool cl, int p, int half, float (&v)[32])
{
    if (cl) {
        const float4 *src = reinterpret_cast<const float4 *>(root) + (half * 8) * kP + p;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // L2-only (the pool / scratch are written by this launch, and allocating these once-read rows in L1 would compete with the
            // tensor core's operand fetch for the shared-memory / L1 data array: measured -1.4 % on the search)
            const float4 q = __ldcg(src + j * kP);
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    } else {
        const float *src = root + (size_t)(half * 32) * kP + p;
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = src[(size_t)c * kP];
    }
}

// 16 consecutive channels [c0, c0 + 16) of pixel p of one root (same layouts)
__device__ __forceinline__ void store_row16(float *root, bool cl, int p, int c0, const float (&v)[16])
{
    if (cl) {
        float4 *dst = reinterpret_cast<float4 *>(root) + (c0 >> 2) * kP + p;
#pragma unroll
        for (int j = 0; j < 4; ++j) __stcg(dst + j * kP, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
    } else {
        float *dst = root + (size_t)c0 * kP + p;
#pragma unroll
        for (int c = 0; c < 16; ++c) dst[(size_t)c * kP] = v[c];
    }
}

// the tree of a warp (tree_persist.cuh) parked in shared memory while the warp does network work: 16 uniform words + 3 x 32
// per-lane words
constexpr int kTreeParkWords = 16 + 3 * 32;
__device__ __forceinline__ void ptree_park(const PTree &T, uint32_t *w, int lane)
{
    if (lane == 0) {
        w[0] = (uint32_t)T.nl; w[1] = (uint32_t)T.plen; w[2] = (uint32_t)T.vtp; w[3] = __float_as_uint(T.mmax); w[4] = __float_as_uint(T.mmin);
        w[5] = (uint32_t)T.root_visit; w[6] = __float_as_uint(T.root_vsum); w[7] = __float_as_uint(T.root_reward);
        w[8] = (uint32_t)T.root_to_play; w[9] = (uint32_t)T.tp0; w[10] = (uint32_t)T.players;
    }
    w[16 + lane] = (uint32_t)T.my_legal; w[48 + lane] = (uint32_t)T.my_pslot; w[80 + lane] = (uint32_t)T.my_pact;
    __syncwarp();
}
__device__ __forceinline__ void ptree_unpark(PTree &T, const uint32_t *w, int lane)
{
    T.nl = (int)w[0]; T.plen = (int)w[1]; T.vtp = (int)w[2]; T.mmax = __uint_as_float(w[3]); T.mmin = __uint_as_float(w[4]);
    T.root_visit = (int)w[5]; T.root_vsum = __uint_as_float(w[6]); T.root_reward = __uint_as_float(w[7]);
    T.root_to_play = (int)w[8]; T.tp0 = (int)w[9]; T.players = (int)w[10];
    T.my_legal = (int)w[16 + lane]; T.my_pslot = (int)w[48 + lane]; T.my_pact = (int)w[80 + lane];
}

__global__ void __launch_bounds__(kTcThreads, 1) k_net_tc(TcNet net, TcIO io, TreeParams tp)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *act = smem;                                   // [2 parts][8 planes][400 rows][16 B]
    unsigned char *ring = smem + kActBytes;                      // [kStages][8 k-groups][128 rows: 64 hi | 64 lo][16 B]
    unsigned char *headw = ring + kStages * kTapBytes;           // [3 heads][hi 2 KB | lo 2 KB]
    TcBars *bars = reinterpret_cast<TcBars *>(headw + kHeadWBytes);
    float *bn_s = reinterpret_cast<float *>(headw + kHeadWBytes + 1024);   // [nlayers][scale 64 | shift 64]

    // warp index as a warp-UNIFORM value (shuffle broadcast): the role branches below become uniform branches, so ptxas keeps the
    // MMA issuer's descriptors in uniform registers instead of moving every operand through R2UR per tcgen05.mma
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int R = io.roots_per_cta;
    const int root0 = blockIdx.x * R;
    const int nvalid = min(R, io.B - root0);                     // roots of this CTA that exist
    // Root groups (io.split_rx > 0): group X = roots 0 .. Rx-1 in tiles 0 .. NTx-1, group Y = the rest from the next tile boundary on.
    // Roots never interact, so the groups move through the layers independently: the MMA warp alternates X.L, Y.L, X.(L+1), ... and
    // the epilogue of one group's layer runs underneath the other group's MMAs (the in-place activation buffer only serialises a
    // group with itself).  The split is only made when it costs no extra M tile (7 roots -> {5, 2}: 2 + 1 tiles).
    const int Rx = (io.split_rx > 0 && io.split_rx < R) ? io.split_rx : R;
    const int NTx = (Rx * kRowsPerRoot + 127) >> 7;
    const int NT = NTx + (((R - Rx) * kRowsPerRoot + 127) >> 7);
    const int ngroups = (Rx < R) ? 2 : 1;
    const RowMap rm = {Rx, R, NTx * 128, nvalid};
    const int npass = io.npass;
    const int nlayers = net.nlayers;
    const int nsims = io.nsims > 0 ? io.nsims : 1;     // > 1 (or persistent): the whole search loop runs inside this launch
    const bool persistent = io.persistent != 0;

    pdl_launch_dependents();      // the next tree kernel may become resident; it blocks in pdl_wait() until this grid is done
    // ---- one-time setup ----
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(&bars->acc_ready[g], 1); mbar_init(&bars->act_ready[g], kEpiThreads); }
        mbar_init(&bars->rew_ready, ngroups);
        mbar_init(&bars->vp_ready, ngroups);
        mbar_init(&bars->fcb_ready, kEpiThreads);
        mbar_init(&bars->fc1_done, 1);
        mbar_init(&bars->fc2b_ready, kEpiThreads);
        mbar_init(&bars->fc2_done, 1);
        fence_mbar_init();
    }
    // heads whose fully connected parts run in this kernel (EfficientZero: the reward features go to the LSTM kernels instead),
    // and the number of 16 KB weight stages they stream through the ring per simulation, after the conv taps: FC1 of all heads
    // as one [128 rows][576] operand (18 stages of 2 k-steps), then one stage per 128-output tile of each head's FC2
    const int hmask_fc = ((net.has_reward && !io.ez_feat) ? 1 : 0) | 6;
    int nfc = kFc1Stages;
    for (int h = 0; h < 3; ++h)
        if ((hmask_fc >> h) & 1) nfc += (net.fc[h].K + 127) >> 7;
    if (warp == kEpiWarps + 1) tmem_alloc(&bars->tmem_base, kTmemCols);
    // 1x1 head weights (12 KB) and the folded BatchNorm tables of this program's layers: plain copies
    for (int i = tid; i < kHeadWBytes / 16; i += kTcThreads)
        reinterpret_cast<uint4 *>(headw)[i] = __ldg(reinterpret_cast<const uint4 *>(net.headw) + i);
    for (int i = tid; i < nlayers * 128; i += kTcThreads)
        bn_s[i] = __ldg(net.bn + (size_t)net.layer_w[i >> 7] * 128 + (i & 127));
    // persistent search: the exploration-rate table of the PUCT rule next to them when it fits (else it is read from global)
    const bool fast_tree = persistent;      // tree_persist.cuh (A <= 32, checked by tc_launch; larger action spaces use the multi-launch graph)
    const float *pbc_tab = tp.pbc;
    if (fast_tree && (nlayers * 128 + tp.N + 1) * 4 <= kBnSmemBytes) {
        float *pbc_s = bn_s + nlayers * 128;
        for (int i = tid; i <= tp.N; i += kTcThreads) pbc_s[i] = tp.pbc[i];
        pbc_tab = pbc_s;
    }
    // the FC2 bias vectors of the heads this kernel evaluates, back to back behind them when they fit as well
    const float *b2_s = nullptr;
    {
        int nb2 = 0;
        for (int h = 0; h < 3; ++h)
            if ((hmask_fc >> h) & 1) nb2 += net.fc[h].K;
        const int used = nlayers * 128 + ((pbc_tab != tp.pbc) ? tp.N + 1 : 0);
        if ((used + nb2) * 4 <= kBnSmemBytes) {
            float *dst = bn_s + used;
            int off = 0;
            for (int h = 0; h < 3; ++h) {
                if (!((hmask_fc >> h) & 1)) continue;
                const Head &H = h == 0 ? net.reward : (h == 1 ? net.value : net.policy);
                for (int i = tid; i < H.K; i += kTcThreads) dst[off + i] = __ldg(H.b2 + i);
                off += H.K;
            }
            b2_s = dst;
        }
    }
    float *ho = reinterpret_cast<float *>(smem + kSmemMain + kHeadScratch);      // tree <-> network hand-off (persistent search)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, bars->tmem_base, 0);      // uniform as well

    if (warp == kEpiWarps) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t n = 0;
            auto stage_in = [&](const unsigned char *src, uint32_t bytes = kTapBytes) {      // every stage is released by a tcgen05.commit of its consumer MMAs
                const int st = n % kStages;
                if (n >= (uint32_t)kStages) mbar_wait(&bars->empty[st], ((n / kStages) - 1) & 1);
                mbar_expect_tx(&bars->full[st], bytes);
                bulk_g2s(ring + st * kTapBytes, src, bytes, &bars->full[st]);
                ++n;
            };
            for (int sim = 0; sim < nsims; ++sim) {     // runs ahead of the consumers: the next simulation's first taps are
                for (int L = 0; L < nlayers; ++L) {     // already in the ring while the tree work is going on
                    const unsigned char *src = net.convw + (size_t)net.layer_w[L] * (9 * kTapBytes);
                    for (int g = 0; g < ngroups; ++g)
                        for (int tap = 0; tap < 9; ++tap) stage_in(src + (size_t)tap * kTapBytes);
                }
                // the heads: FC1 weights of all heads (18 stages), then the FC2 tiles of every head this kernel evaluates
                for (int i = 0; i < kFc1Stages; ++i) stage_in(net.fcw + (size_t)i * kFc1StageBytes, kFc1StageBytes);
                for (int h = 0; h < 3; ++h) {
                    if (!((hmask_fc >> h) & 1)) continue;
                    const int nblk = (net.fc[h].K + 127) >> 7;
                    for (int blk = 0; blk < nblk; ++blk) stage_in(net.fcw + net.fc[h].fc2_off + (size_t)blk * kTapBytes);
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ================= MMA issuer: the whole warp runs the loops (uniform control flow), one elected lane issues =================
        const uint32_t act_s = smem_u32(act), ring_s = smem_u32(ring), headw_s = smem_u32(headw);
        const uint32_t idesc128 = make_idesc_f16(128, 128), idesc64 = make_idesc_f16(128, 64);
        const uint32_t idesc16 = make_idesc_f16(128, 16), idesc32 = make_idesc_f16(128, 32);
        const uint32_t a_lbo = kPlaneBytes >> 4, a_sbo = 8;
        const uint64_t a_desc0 = make_desc(act_s + kMargin * 16, a_lbo, a_sbo);   // row 0, hi part, k-step 0
        const uint64_t b_desc0 = make_desc(ring_s, kTapKgBytes >> 4, 8);          // stage 0, k-group 0: rows 0-63 hi, 64-127 lo
        unsigned long long *dbg = (io.dbg && blockIdx.x == 0) ? io.dbg : nullptr;
        uint32_t n = 0;
        // 1x1 head convolutions on layer L's OUTPUT for the tiles [t0, t1) of one root group; the caller has waited for that group's
        // epilogue.  The value/policy result reuses drained conv accumulator columns (tile 0's), so that hook is only legal on the LAST
        // layer; the reward result has its own columns.  One arrival per group on the hook's barrier.
        auto hooks = [&](int L, int t0, int t1) {
            const int flags = net.layer_flags[L];
            for (int hook = 0; hook < 2; ++hook) {
                if (!(flags & (hook == 0 ? LF_HOOK_REWARD : LF_HOOK_VALPOL))) continue;
                for (int t = t0; t < t1; ++t) {
                    const uint32_t arow = act_s + (uint32_t)(kMargin + t * 128) * 16u;
                    for (int ps = 0; ps < npass; ++ps) {
                        const uint32_t apart = (ps == 2) ? kPartBytes : 0;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t ad = make_desc(arow + apart + ks * 2 * kPlaneBytes, a_lbo, a_sbo);
                            if (hook == 0) {   // reward head: N = 16, [kg][16 co][8]
                                const uint32_t wb = headw_s + ((ps == 1) ? 2048 : 0);
                                umma_f16_elect(tmem + kColRew + t * 16, ad, make_desc(wb + ks * 512, 16, 8), idesc16, (ps | ks) != 0);
                            } else {           // value + policy heads as one N = 32 matrix, [kg][32 co][8]
                                const uint32_t wb = headw_s + 4096 + ((ps == 1) ? 4096 : 0);
                                umma_f16_elect(tmem + kColVp + t * 32, ad, make_desc(wb + ks * 1024, 32, 8), idesc32, (ps | ks) != 0);
                            }
                        }
                    }
                }
                umma_commit_elect(hook == 0 ? &bars->rew_ready : &bars->vp_ready);
            }
        };
        for (int sim = 0; sim < nsims; ++sim) {
        for (int L = 0; L < nlayers; ++L) {
            const uint32_t ev = (uint32_t)sim * (nlayers + 1) + L;   // act_ready event index: 1 load + nlayers epilogues per simulation
            for (int g = 0; g < ngroups; ++g) {
                const int t0 = g == 0 ? 0 : NTx, t1 = g == 0 ? NTx : NT;
                mbar_wait_converged(&bars->act_ready[g], ev & 1);    // this group's inputs written, its TMEM accumulators drained
                tc_fence_after();
                if (L > 0) hooks(L - 1, t0, t1);                     // head convolutions on the previous layer's output
                if (dbg && g == 0) dbg[32 + 2 * L] = clock64();
                for (int ti = 0; ti < 9; ++ti) {
                    const int tap = ti;
                    const uint32_t pos = n++;
                    const int st = pos % kStages;
                    const long long tw0 = dbg ? clock64() : 0;
                    mbar_wait_converged(&bars->full[st], (pos / kStages) & 1);
                    if (dbg) dbg[52 + (L == 0 && tap == 0 && g == 0 ? 1 : 0)] += (unsigned long long)(clock64() - tw0);   // [52] ring waits, [53] first tap of a simulation
                    tc_fence_after();
                    const int shift = (tap / 3 - 1) * kPitch + (tap % 3 - 1);
                    // descriptors differ only in the 14-bit start-address field: add 16-byte-unit offsets to a base
                    const uint64_t b0 = b_desc0 + (uint64_t)((st * kTapBytes) >> 4);
                    // fp32-accurate mode, per (tile, tap, k-step): A_hi x [B_hi | B_lo] as ONE N = 128 MMA (the tap block holds the two
                    // weight parts as 128 operand rows; the two 64-column halves are added at read-out), then A_lo x B_hi (N = 64, into
                    // columns 0-63).  Measured alternatives (three N = 64 MMAs; N = 128 twice; all tiles' N = 128 first) were slower.
                    const uint32_t kA = (2 * kPlaneBytes) >> 4, kB = (2 * kTapKgBytes) >> 4, kLo = kPartBytes >> 4;
                    for (int t = t0; t < t1; ++t) {
                        const uint64_t a0 = a_desc0 + (uint64_t)(t * 128 + shift);
                        const uint32_t d = tmem + kColAcc + t * kAccCols;
                        if (npass != 3) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_elect(d, a0 + ks * kA, b0 + ks * kB, idesc64, (ti | ks) != 0);
                        } else {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_elect(d, a0 + ks * kA, b0 + ks * kB, idesc128, (ti | ks) != 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_elect(d, a0 + kLo + ks * kA, b0 + ks * kB, idesc64, 1);
                        }
                    }
                    umma_commit_elect(&bars->empty[st]);               // frees this ring slot when the MMAs have read it
                }
                umma_commit_elect(&bars->acc_ready[g]);
                if (dbg && g == ngroups - 1) dbg[33 + 2 * L] = clock64();
            }
        }
        // head convolutions on the last layer's output (waiting twice on a completed phase is immediate)
        for (int g = 0; g < ngroups; ++g) {
            mbar_wait_converged(&bars->act_ready[g], ((uint32_t)sim * (nlayers + 1) + nlayers) & 1);
            tc_fence_after();
            hooks(nlayers - 1, g == 0 ? 0 : NTx, g == 0 ? NTx : NT);
        }
        // ---- the heads' fully connected layers (weights = M operand from the ring, roots = N columns; see the heads section)
        {
            const uint32_t fb_s = act_s, hb_s = act_s + kFbBytes;
            const uint32_t idesc_fc1 = make_idesc_f16(128, 64), idesc_fc2 = make_idesc_f16(128, 16);
            mbar_wait_converged(&bars->fcb_ready, sim & 1);          // features of all heads are in shared memory
            tc_fence_after();
            for (int i = 0; i < kFc1Stages; ++i, ++n) {
                const int st = n % kStages;
                mbar_wait_converged(&bars->full[st], (n / kStages) & 1);
                tc_fence_after();
#pragma unroll
                for (int ksi = 0; ksi < 2; ++ksi) {
                    const int kstep = 2 * i + ksi;
                    // [k-step][hi 3 KB | lo 3 KB], [kg 2][96 rows][8]: rows 96-127 of the M = 128 operand are whatever follows (unused lanes)
                    const uint64_t a_hi = make_desc(ring_s + st * kTapBytes + ksi * (kFc1StageBytes / 2), 1536 >> 4, 8), a_lo = a_hi + (3072 >> 4);
                    const uint64_t b = make_desc(fb_s + kstep * 2 * kFbKgBytes, kFbKgBytes >> 4, 8);
                    umma_f16_elect(tmem + kColFc1, a_hi, b, idesc_fc1, kstep != 0);
                    umma_f16_elect(tmem + kColFc1, a_lo, b, idesc_fc1, 1);
                }
                umma_commit_elect(&bars->empty[st]);
            }
            umma_commit_elect(&bars->fc1_done);
            mbar_wait_converged(&bars->fc2b_ready, sim & 1);         // hidden activations (FC2's B operand) written
            tc_fence_after();
            int tile = 0;
            for (int h = 0; h < 3; ++h) {
                if (!((hmask_fc >> h) & 1)) continue;
                const int nblk = (net.fc[h].K + 127) >> 7;
                for (int blk = 0; blk < nblk; ++blk, ++tile, ++n) {
                    const int st = n % kStages;
                    mbar_wait_converged(&bars->full[st], (n / kStages) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t a_hi = make_desc(ring_s + st * kTapBytes + ks * 2 * 2048, 2048 >> 4, 8), a_lo = a_hi + (8192 >> 4);
                        const uint64_t b = make_desc(hb_s + h * kHbHeadBytes + ks * 2 * 256, 256 >> 4, 8);
                        umma_f16_elect(tmem + kColFc2 + tile * 16, a_hi, b, idesc_fc2, ks != 0);
                        umma_f16_elect(tmem + kColFc2 + tile * 16, a_lo, b, idesc_fc2, 1);
                    }
                    umma_commit_elect(&bars->empty[st]);
                }
            }
            umma_commit_elect(&bars->fc2_done);
        }
        }
    } else {
        // ================= epilogue warps: warp w owns TMEM lanes 32*(w%4).. and the 32-column half w/4 =================
        const int q4 = warp & 3, half = warp >> 2, rowid = q4 * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
        unsigned char *fr = smem + kSmemMain;      // reward features (fp16 hi / lo, FC1's B-operand rows 0-7), parked from their hook
        unsigned long long *dbg = (io.dbg && blockIdx.x == 0 && tid == 0) ? io.dbg : nullptr;
        if (dbg) dbg[0] = clock64();
        pdl_wait();                   // ix / action / the latent pool come from the preceding kernels
        int acc_par = 0;
        // this warp's tree (tree_persist.cuh): its scalars / first path entries live in registers during the tree phase and are
        // parked in shared memory while the warp does network work
        uint32_t *tree_park = reinterpret_cast<uint32_t *>(smem + kSmemMain + kFrBytes) + warp * kTreeParkWords;
        if (fast_tree && warp < nvalid) {
            PTree T;
            ptree_init(tp, T, root0 + warp, lane);
            ptree_park(T, tree_park, lane);
        }
        // row bookkeeping of this thread (simulation-invariant): per tile (root slot in the CTA) << 8 | pixel, or -1 for pad rows
        int rowc[kMaxTiles];
#pragma unroll
        for (int t = 0; t < kMaxTiles; ++t) {
            int r, p;
            const bool valid = row_decode(rm, t * 128 + rowid, r, p) && (t < NT);
            rowc[t] = valid ? ((r << 8) | p) : -1;
        }
        for (int sim = 0; sim <= nsims; ++sim) {        // iteration nsims: only the back-up of the last simulation (mcts_ctree.py:365-368)
            if (sim == nsims && !persistent) break;
            if (dbg && sim < nsims) dbg[50] = clock64();
            if (persistent) {
                // ---- tree phase: one warp per root of this CTA (roots never interact, so the whole search of these roots
                // lives in this CTA): back up the previous simulation, then descend to the next leaf.  cnode.cpp:480-500,754-825
                if (warp < nvalid) {
                    const int b = root0 + warp;
                    PTree T;
                    ptree_unpark(T, tree_park, lane);
                    float *my_ho = ho + warp * kHoWords;
                    if (sim > 0)       // network outputs of the previous simulation: from the hand-off (written by this CTA's read-out)
                        ptree_backprop(tp, T, b, lane, io.sim0 + sim, my_ho[3], my_ho[2], my_ho + 4);
                    if (dbg && sim < nsims) dbg[56] = clock64();
                    if (sim < nsims) {
                        ptree_traverse(tp, T, b, lane, io.deterministic, (unsigned)(io.sim0 + sim), pbc_tab, io.ix_rw, io.action_rw,
                                       reinterpret_cast<int *>(my_ho), reinterpret_cast<int *>(my_ho) + 1);
                        ptree_park(T, tree_park, lane);
                    }
                }
                if (sim == nsims) break;
                __threadfence_block();
                asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
            }
            if (dbg) dbg[51] = clock64();
            float *latent_out = persistent ? (io.latent_pool_rw + (size_t)(io.sim0 + sim + 1) * io.slot_stride) : io.latent_out;
            // zero the margins (the head scratch of the previous simulation overlays them; pad rows inside the tiles are
            // rewritten as zeros by every load / epilogue)
            for (int i = tid; i < 2 * 8 * 2 * kMargin; i += kEpiThreads) {
                int part = i / (8 * 2 * kMargin), rem = i % (8 * 2 * kMargin), plane = rem / (2 * kMargin), r = rem % (2 * kMargin);
                int row = r < kMargin ? r : kMargin + kMaxTiles * 128 + (r - kMargin);
                *reinterpret_cast<uint4 *>(act + part * kPartBytes + plane * kPlaneBytes + row * 16) = make_uint4(0, 0, 0, 0);
            }
            // the pool slot holding the input latent of each of this thread's rows (tree -> network hand-off)
            int slot[kMaxTiles];
#pragma unroll
            for (int t = 0; t < kMaxTiles; ++t)
                slot[t] = rowc[t] < 0 ? 0 : (persistent ? reinterpret_cast<const int *>(ho)[(rowc[t] >> 8) * kHoWords] : (io.ix ? io.ix[root0 + (rowc[t] >> 8)] : 0));
            auto in_ptr = [&](int t) { return io.latent_base + (size_t)slot[t] * io.slot_stride + (size_t)(root0 + (rowc[t] >> 8)) * (kC * kP); };
            auto in_is_cl = [&](int t) { return io.pool_cl != 0 && slot[t] > 0; };      // slot 0: root latents as the API delivered them (NCHW)
            // ---- load the input activation: gather the latents (two tiles' loads in flight), split to fp16 hi/lo ----
            {
                float va[32], vb[32];
                auto fetch_in = [&](int t, float (&dst)[32]) {
                    if (rowc[t] >= 0) load_row32(in_ptr(t), in_is_cl(t), rowc[t] & 255, half, dst);
                    else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) dst[c] = 0.0f;
                    }
                };
                auto put = [&](int t, const float (&src)[32]) {
                    const int m = t * 128 + rowid;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        unsigned char *p = act + (half * 4 + g) * kPlaneBytes + (kMargin + m) * 16;
                        store_split8(p, p + kPartBytes, src + 8 * g);
                    }
                };
                // phase 0 of a root group's act_ready: its layer 0 may start (group X while group Y's rows are still being loaded)
                auto loaded = [&](int t) {
                    if (t == NTx - 1 || t == NT - 1) {
                        fence_proxy_async();
                        tc_fence_before();
                        mbar_arrive(&bars->act_ready[t == NT - 1 ? ngroups - 1 : 0]);
                    }
                };
                fetch_in(0, va);
                if (NT > 1) fetch_in(1, vb);
                put(0, va);
                loaded(0);
                if (NT > 2) fetch_in(2, va);
                if (NT > 1) { put(1, vb); loaded(1); }
                if (NT > 2) { put(2, va); loaded(2); }
            }
            if (dbg) dbg[1] = clock64();

            bool skip_in_scratch = false;      // the residual operand: the input latent until a layer has parked its output
            bool pending_rew = false;
            auto reward_features = [&]() {
                    // reward 1x1 accumulators -> BN/ReLU features, parked (fp16 hi / lo) until the heads' FC pass at the end of the
                    // simulation -- or, EfficientZero (efficientzero_model.py:556-562), written out as the input of the LSTM that the
                    // next kernel evaluates as one batched GEMM over all roots (ez.cu).  Runs underneath the NEXT layer's MMAs.
                    mbar_wait_warp(&bars->rew_ready, sim & 1);
                    tc_fence_after();
                    if (io.ez_feat) {
                        if (half == 0) {
                            const int nin = net.hc[0] * kP;
#pragma unroll
                            for (int t = 0; t < kMaxTiles; ++t) {
                                if (t >= NT) continue;
                                float v[16];
                                tmem_ld16(lane_base + kColRew + t * 16, v);
                                if (rowc[t] >= 0)
                                    for (int c = 0; c < net.hc[0]; ++c)
                                        io.ez_feat[(size_t)(root0 + (rowc[t] >> 8)) * nin + c * kP + (rowc[t] & 255)] =
                                            fmaxf(fmaf(v[c], __ldg(net.head_bn + c), __ldg(net.head_bn + 16 + c)), 0.0f);
                            }
                        }
                        tc_fence_before();
                    } else {
                        head_scatter(net, 1, fr, nullptr, tmem, NT, rm);
                    }
                    if (dbg) dbg[28] = clock64();
            };
            for (int L = 0; L < nlayers; ++L) {
                const int flags = net.layer_flags[L];
                const float4 *bn4 = reinterpret_cast<const float4 *>(bn_s + L * 128 + half * 32);   // scale; shift 16 float4 further
                const bool park = (flags & LF_STORE_RES) && (L + 1 < nlayers);
                // Residual operands (and, for the dynamics conv, the action-plane bias) are thread-private global rows (L2-resident).
                // Two row buffers stay in flight so that no load latency is exposed after the accumulators arrive: tiles 0 and 1 are
                // fetched BEFORE the accumulator wait (this thread idles through the layer's MMAs anyway), tile 2's skip right after
                // tile 0 is done and its action bias right after tile 1.
                const bool has_res = (flags & LF_RES) != 0, has_ab = (flags & LF_ACT_BIAS) != 0;
                float ra[32], rb[32];
                auto fetch_skip = [&](int t, float (&dst)[32]) {
                    if (rowc[t] < 0) return;
                    if (skip_in_scratch) load_row32(io.skip_scratch + (size_t)(root0 + (rowc[t] >> 8)) * (kC * kP), true, rowc[t] & 255, half, dst);
                    else load_row32(in_ptr(t), in_is_cl(t), rowc[t] & 255, half, dst);
                };
                auto fetch_abias = [&](int t, float (&dst)[32], bool add) {
                    if (rowc[t] < 0) return;
                    const int action_raw = persistent ? reinterpret_cast<const int *>(ho)[(rowc[t] >> 8) * kHoWords + 1] : io.action[root0 + (rowc[t] >> 8)];
                    const int action = min(max(action_raw, 0), net.A - 1);
                    const float4 *ab = reinterpret_cast<const float4 *>(net.abias) + ((size_t)action * 16 + half * 8) * kP + (rowc[t] & 255);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 q = __ldcg(ab + j * kP);      // 83 KB table, one row per (root, action): L2-only like the skip rows
                        if (add) { dst[4 * j] += q.x; dst[4 * j + 1] += q.y; dst[4 * j + 2] += q.z; dst[4 * j + 3] += q.w; }
                        else { dst[4 * j] = q.x; dst[4 * j + 1] = q.y; dst[4 * j + 2] = q.z; dst[4 * j + 3] = q.w; }
                    }
                };
                if (has_res) {
                    fetch_skip(0, ra);
                    if (NT > 1) fetch_skip(1, rb);
                    if (has_ab) {
                        fetch_abias(0, ra, true);
                        if (NT > 1) fetch_abias(1, rb, true);
                    }
                }
                mbar_wait_warp(&bars->acc_ready[0], acc_par);
                tc_fence_after();
                if (dbg) dbg[2 + 2 * L] = clock64();
#pragma unroll
                for (int t = 0; t < kMaxTiles; ++t) {
                    if (t >= NT) continue;
                    if (t > 0 && t == NTx) {
                        // root group X is complete: its next layer (or hook) may start while group Y's tiles are processed
                        fence_proxy_async();
                        tc_fence_before();
                        mbar_arrive(&bars->act_ready[0]);
                        if (dbg) dbg[57] = clock64();
                        mbar_wait_warp(&bars->acc_ready[1], acc_par);
                        tc_fence_after();
                    }
                    const int m = t * 128 + rowid;
                    const bool valid = rowc[t] >= 0;
                    const int p = rowc[t] & 255, b = root0 + (rowc[t] >> 8);
                    float (&rs)[32] = (t == 1) ? rb : ra;
                    if (has_res && has_ab && t == 2) {
                        // tile 2's action bias arrived in the other buffer: the same (skip + bias) association as tiles 0 / 1, so
                        // that a root's bits do not depend on which tile of the CTA it lands in
#pragma unroll
                        for (int c = 0; c < 32; ++c) ra[c] += rb[c];
                    }
                    // the thread's 32 channels in two halves of 16 (keeps the live registers of this loop under the budget)
#pragma unroll
                    for (int hs = 0; hs < 2; ++hs) {
                        uint32_t ua[16];
                        float v[16];
                        tmem_ld16_issue(lane_base + kColAcc + t * kAccCols + half * 32 + hs * 16, ua);
                        if (npass == 3) {
                            uint32_t ub[16];
                            tmem_ld16_issue(lane_base + kColAcc + t * kAccCols + 64 + half * 32 + hs * 16, ub);
                            tmem_ld_wait();
                            tmem_pin(ua); tmem_pin(ub);
#pragma unroll
                            for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(ua[c]) + __uint_as_float(ub[c]);
                        } else {
                            tmem_ld_wait();
                            tmem_pin(ua);
#pragma unroll
                            for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(ua[c]);
                        }
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 sc = bn4[hs * 4 + j], sh = bn4[16 + hs * 4 + j];
                                v[4 * j] = fmaf(v[4 * j], sc.x, sh.x); v[4 * j + 1] = fmaf(v[4 * j + 1], sc.y, sh.y);
                                v[4 * j + 2] = fmaf(v[4 * j + 2], sc.z, sh.z); v[4 * j + 3] = fmaf(v[4 * j + 3], sc.w, sh.w);
                            }
                            if (has_res) {
#pragma unroll
                                for (int c = 0; c < 16; ++c) v[c] += rs[hs * 16 + c];
                            }
#pragma unroll
                            for (int c = 0; c < 16; ++c) v[c] = fmaxf(v[c], 0.0f);       // every layer of these programs ends in ReLU
                            if (park) store_row16(io.skip_scratch + (size_t)b * (kC * kP), true, p, half * 32 + hs * 16, v);
                            if (flags & LF_WRITE_LATENT) {
                                if (latent_out) store_row16(latent_out + (size_t)b * (kC * kP), io.pool_cl != 0, p, half * 32 + hs * 16, v);
                                if (io.latent_out2) store_row16(io.latent_out2 + (size_t)b * (kC * kP), false, p, half * 32 + hs * 16, v);
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 16; ++c) v[c] = 0.0f;          // pad rows / absent roots: the conv padding of the next layer
                        }
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            unsigned char *pp = act + (half * 4 + hs * 2 + g) * kPlaneBytes + (kMargin + m) * 16;
                            store_split8_pos(pp, pp + kPartBytes, v + 8 * g);
                        }
                    }
                    // next fetches into the buffer this tile has just released
                    if (has_res && t == 0 && NT > 2) fetch_skip(2, ra);
                    if (has_res && has_ab && t == 1 && NT > 2) fetch_abias(2, rb, false);
                }
                if (park) skip_in_scratch = true;
                acc_par ^= 1;                                       // one commit per layer and group, across simulations
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&bars->act_ready[ngroups - 1]);         // phase L+1: next layer / this layer's hook may start
                if (dbg) dbg[3 + 2 * L] = clock64();
                // The reward hook's 1x1 accumulators are consumed one layer LATER: group Y's hook MMAs are only issued at the start of its
                // next job (after group X's next-layer MMAs), and waiting for them here would hold group X's epilogue back by a whole job
                if (pending_rew) { reward_features(); pending_rew = false; }
                if ((flags & LF_HOOK_REWARD) && net.has_reward) pending_rew = true;
            }
            if (pending_rew) { reward_features(); pending_rew = false; }

            // the value / policy 1x1 accumulators must be complete before the head stage reads them / reuses the buffer
            mbar_wait_warp(&bars->vp_ready, sim & 1);
            tc_fence_after();
            if (dbg) dbg[24] = clock64();
            // ---- heads: 1x1 accumulators -> BN/ReLU features, then FC1 -> FC2 -> softmax expectation -> h^-1 for all heads with the
            // weights streamed through the ring.  The scratch overlays the activation buffer: every conv MMA of this simulation
            // has completed (ordered through act_ready -> MMA issuer -> vp_ready; the explicit barrier keeps racecheck, which
            // does not follow mbarriers, quiet)
            {
                unsigned char *fb = act, *hb = act + kFbBytes;
                float *red = reinterpret_cast<float *>(hb + 3 * kHbHeadBytes);
                asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
                if (hmask_fc & 1) {      // the parked reward features become rows 0-7 (hi) / 32-39 (lo) of the B operand
                    for (int i = tid; i < 2 * 72 * 8; i += kEpiThreads) {
                        const int part = i / (72 * 8), rem = i - part * (72 * 8), kg = rem >> 3, r = rem & 7;
                        *reinterpret_cast<uint4 *>(fb + kg * kFbKgBytes + (part * 32 + r) * 16) =
                            *reinterpret_cast<const uint4 *>(fr + part * (kFrBytes / 2) + (kg * 8 + r) * 16);
                    }
                }
                head_scatter(net, 6, nullptr, fb, tmem, NT, rm);
                fence_proxy_async();
                mbar_arrive(&bars->fcb_ready);                      // FC1 may start
                if (dbg) dbg[44] = clock64();
                mbar_wait_warp(&bars->fc1_done, sim & 1);
                tc_fence_after();
                if (dbg) dbg[45] = clock64();
                heads_hidden(net, hmask_fc, hb, tmem);
                fence_proxy_async();
                mbar_arrive(&bars->fc2b_ready);                     // FC2 may start
                if (dbg) dbg[46] = clock64();
                mbar_wait_warp(&bars->fc2_done, sim & 1);
                tc_fence_after();
                if (dbg) dbg[47] = clock64();
                heads_outputs(net, io, hmask_fc, tmem, red, nvalid, root0, b2_s, persistent ? ho : nullptr);
            }
            if (dbg) dbg[27] = clock64();
            __threadfence_block();
            asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) {
        __syncwarp();
        tmem_dealloc(tmem, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------- host
static void split_half(float v, float scale, __half &hi, __half &lo)
{
    float s = v * scale;
    hi = __float2half_rn(s);
    lo = __float2half_rn(s - __half2float(hi));
}

// One 3x3 conv -> 9 tap blocks, each [kg = ci/8][co: 64 hi rows | 64 lo rows][ci % 8] fp16.  Returns the power-of-two
// scale applied to the weights (exact), which the caller folds into the BatchNorm scale.
float tc_pack_conv3(const float *w_torch /*[64][cin][3][3]*/, int cin_total, int cin_used, unsigned char *dst)
{
    float mx = 0.0f;
    for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < cin_used; ++ci)
            for (int t = 0; t < 9; ++t) mx = std::max(mx, fabsf(w_torch[((size_t)co * cin_total + ci) * 9 + t]));
    int e = 0;
    if (mx > 0.0f) frexpf(mx, &e);               // mx = f * 2^e, f in [0.5, 1)
    const float scale = ldexpf(1.0f, 13 - e);    // largest |w| lands in [4096, 8192)
    __half *h = reinterpret_cast<__half *>(dst);
    for (int t = 0; t < 9; ++t)
        for (int co = 0; co < 64; ++co)
            for (int ci = 0; ci < 64; ++ci) {
                __half hi, lo;
                split_half(w_torch[((size_t)co * cin_total + ci) * 9 + t], scale, hi, lo);
                const size_t off = (size_t)t * (kTapBytes / 2) + ((size_t)(ci / 8) * 128 + co) * 8 + (ci % 8);
                h[off] = hi;
                h[off + 64 * 8] = lo;             // rows 64-127 of the k-group: the lo parts ([B_hi | B_lo] is one N = 128 operand)
            }
    return scale;
}

// 1x1 head conv [hc][64] -> [hi | lo] with nco rows (zero padded), [kg][co][8].
float tc_pack_conv1(const float *w /*[hc][64]*/, int hc, int nco, int co_offset, unsigned char *dst_hi, unsigned char *dst_lo)
{
    float mx = 0.0f;
    for (int i = 0; i < hc * 64; ++i) mx = std::max(mx, fabsf(w[i]));
    int e = 0;
    if (mx > 0.0f) frexpf(mx, &e);
    const float scale = ldexpf(1.0f, 13 - e);
    __half *hh = reinterpret_cast<__half *>(dst_hi), *hl = reinterpret_cast<__half *>(dst_lo);
    for (int co = 0; co < hc; ++co)
        for (int ci = 0; ci < 64; ++ci) {
            __half hi, lo;
            split_half(w[co * 64 + ci], scale, hi, lo);
            const size_t off = ((size_t)(ci / 8) * nco + co + co_offset) * 8 + (ci % 8);
            hh[off] = hi;
            hl[off] = lo;
        }
    return scale;
}

int tc_head_layout_bytes() { return kHeadWBytes; }
int tc_conv_layout_bytes() { return 9 * kTapBytes; }

static unsigned long long *g_dbg = nullptr;     // bring-up instrumentation only (env LZ_TC_DEBUG at model finalize time)
unsigned long long *tc_debug_buffer() { return g_dbg; }

int tc_prepare_launch()
{
    LZ_CUDA_CHECK(cudaFuncSetAttribute(k_net_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    if (getenv("LZ_TC_DEBUG") && !g_dbg) {      // allocated here (model finalize), never inside a stream capture
        LZ_CUDA_CHECK(cudaMalloc(&g_dbg, 64 * 8));
        LZ_CUDA_CHECK(cudaMemset(g_dbg, 0, 64 * 8));
    }
    return LZ_OK;
}

int tc_pick_roots(int B)
{
    int r = (B + 147) / 148;
    return std::min(std::max(r, 1), kMaxRoots);
}

// Root-group split of a CTA's R roots (see the kernel): only where {Rx, R - Rx} needs no more 128-row tiles than R roots together.
static int tc_pick_split(int R)
{
    auto tiles = [](int r) { return (r * kRowsPerRoot + 127) / 128; };
    const int cand[3] = {5, 4, 2};
    for (int rx : cand)
        if (rx < R && tiles(rx) + tiles(R - rx) == tiles(R)) return rx;
    return 0;
}

int tc_launch(const TcNet &net, const TcIO &io_in, cudaStream_t s, const TreeParams *tp_in)
{
    TcIO io = io_in;
    TreeParams tp;
    memset(&tp, 0, sizeof(tp));
    if (tp_in) tp = *tp_in;
    LZ_REQUIRE(!io.persistent || tp_in, LZ_EINVAL, "tc_launch: persistent search needs tree parameters");
    LZ_REQUIRE(net.A <= 1000, LZ_EINVAL, "tc_launch: action space %d too large for the head scratch of the tcgen05 path", net.A);
    LZ_REQUIRE(io.skip_scratch, LZ_EINVAL, "tc_launch: no skip scratch");
    io.dbg = g_dbg;
    io.roots_per_cta = tc_pick_roots(io.B);
    LZ_REQUIRE(!io.persistent || tp.A <= 32, LZ_EINVAL, "tc_launch: the persistent search needs A <= 32 (got %d)", tp.A);
    if (const char *e = getenv("LZ_TC_ROOTS")) io.roots_per_cta = std::min(std::max(atoi(e), 1), kMaxRoots);
    io.split_rx = tc_pick_split(io.roots_per_cta);
    if (const char *e = getenv("LZ_TC_SPLIT")) io.split_rx = atoi(e);          // A/B switch (0 = one group)
    const int grid = (io.B + io.roots_per_cta - 1) / io.roots_per_cta;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = io_in.pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    count_launch();
    LZ_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_net_tc, net, io, tp));
    return LZ_OK;
}

}  // namespace lz
