// net_tc.cu -- tcgen05 (5th-gen tensor core) path of the MuZero latent-grid networks for sm_100a.
//
// One CTA runs the WHOLE recurrent_inference (or the latent-grid tail of initial_inference) for up to 7
// roots: five 3x3 convolutions + three 1x1 head convolutions as tcgen05.mma with fp32 accumulators in
// TMEM, BatchNorm/residual/ReLU epilogues out of TMEM, and the small fully connected head layers +
// softmax-expectation + inverse scalar transform on the CUDA cores.  Activations never leave the SM.
//
// fp32 accuracy on fp16 tensor cores ("3xFP16"): every fp32 operand v is split v = hi + lo with
// hi = fp16(v), lo = fp16(v - hi) (22 significant bits); D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi with fp32
// accumulation drops only the 2^-22 lo*lo term.  Same bytes per element as fp32 (2+2), three MMAs at
// the fp16 rate.  Weights are pre-split on the host and pre-scaled by a power of two (exact; folded
// back into the BatchNorm scale) so their lo parts stay in fp16's normal range.  Mode 2 ("fast")
// issues only the hi*hi pass.
//
// Implicit GEMM without im2col: activations live in shared memory as [k-group of 8 channels][row][8
// halves] (the UMMA K-major no-swizzle canonical layout with SBO = 128 B, so row r of the operand is at
// start + 16*r bytes).  Rows are the pixels of a 7-wide padded grid (49 rows per root, column 6 and row
// 6 zero), so the input of output row m for tap (dy,dx) is row m + 7*dy + dx: each of the 9 taps is the
// SAME buffer addressed through a descriptor whose start address is shifted by (7*dy+dx)*16 bytes.  The
// zero pad rows double as the conv padding between rows and between consecutive roots.  M tiles of 128
// rows cut anywhere (every output row only depends on shifted input rows); all tiles of a layer
// accumulate in TMEM before the epilogue rewrites the buffer IN PLACE; the ResBlock skip tensor is parked
// in spare TMEM columns (tcgen05.st) instead of a second shared-memory buffer.
//
// Warp roles (320 threads): warps 0-7 = epilogue / loads / heads (warp w owns TMEM lanes 32*(w%4).. and the
// 32-column half w/4 of every accumulator), warp 8 lane 0 = weight producer (cp.async.bulk global->shared
// ring, mbarrier complete_tx), warp 9 lane 0 = MMA issuer (+ TMEM alloc/dealloc by warp 9).
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "model.cuh"
#include "net_tc.cuh"
#include "tc_ptx.cuh"
#include "tree.cuh"

namespace lz {

// Build with -DLZ_UNIFORM_ISSUE (env LZ_UNIFORM_ISSUE=1 for lightzero_b200/_build.py) to issue the MMAs from uniform control
// flow (all 32 lanes of the issuing warp run the loop, elect.sync picks the issuer).  NOT the default in round 1: the
// measurement that motivates it (profiles/r01e_mma_probe.md) arrived when no GPU time was left to validate the kernel.
#ifdef LZ_UNIFORM_ISSUE
#define LZ_MMA_ISSUER_ON true
#define LZ_UMMA umma_f16_elect
#define LZ_UCOMMIT umma_commit_elect
#else
#define LZ_MMA_ISSUER_ON (lane == 0)
#define LZ_UMMA umma_f16
#define LZ_UCOMMIT umma_commit
#endif

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
constexpr int kTapBytes = 2 * 64 * 64 * 2;       // one 3x3 tap, hi + lo: 16384 B
constexpr int kStages = 4;
constexpr int kHeadWBytes = 3 * 2 * 16 * 64 * 2; // three 1x1 heads (hc <= 16), hi + lo: 12288 B
constexpr int kSmemMain = kActBytes + kStages * kTapBytes + kHeadWBytes + 1024;
constexpr int kHeadScratch = (kMaxRoots * (576 + 32 + 608) + 8 * kMaxRoots * 32) * 4;   // one head: features, hidden, logits, partials
constexpr int kSmemBytes = kSmemMain + kHeadScratch;

// TMEM columns
constexpr int kColAcc = 0;        // 3 tiles x 64
constexpr int kColRes = 192;      // 3 tiles x 64  (ResBlock skip tensor, fp32)
constexpr int kColRew = 384;      // 3 tiles x 16  (reward 1x1)
constexpr int kTmemCols = 512;

struct TcBars {
    uint64_t full[kStages], empty[kStages];
    uint64_t acc_ready;     // MMA -> epilogue: this layer's accumulators are complete
    uint64_t act_ready;     // epilogue -> MMA: activations (and TMEM) are ready for the next layer
    uint64_t rew_ready;     // MMA -> heads: reward 1x1 accumulators complete (single phase)
    uint64_t vp_ready;      // MMA -> heads: value/policy 1x1 accumulators complete (single phase)
    uint32_t tmem_base;
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------- heads
// 1x1-conv accumulators (TMEM) -> BN/ReLU features -> FC1 -> BN/ReLU -> FC2 -> softmax expectation -> h^-1 for the
// heads in `hmask` (bit 0 reward, 1 value, 2 policy).  Executed by the kEpiThreads epilogue threads together
// (named barrier 1).  scr: [nh][7][576] features | [nh][7][32] hidden | [nh][7][ldl] logits | [8][nh][7][32] partials.
__device__ __forceinline__ void head_stage(const TcNet &net, const TcIO &io, int hmask, float *scr, uint32_t tmem, int NT,
                                           int rows_used, int nvalid, int root0)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, half = warp >> 2, rowid = q4 * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    const int A = net.A, ldl = max(608, (A + 31) & ~31);
    const int nh = __popc(hmask);
    int slot[3];
    slot[0] = 0; slot[1] = hmask & 1; slot[2] = (hmask & 1) + ((hmask >> 1) & 1);
    float *hflat = scr, *hidden = hflat + nh * kMaxRoots * 576, *logits = hidden + nh * kMaxRoots * 32;
    float *part = logits + nh * kMaxRoots * ldl;
    // The scratch may overlay the activation buffer the epilogue has just written (value/policy stage).  Those writes are
    // already ordered before this point through act_ready -> MMA issuer -> vp_ready, but an explicit barrier among the
    // epilogue threads costs nothing and keeps compute-sanitizer's racecheck (which does not follow mbarriers) quiet.
    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
    for (int i = tid; i < nh * kMaxRoots * 576; i += kEpiThreads) hflat[i] = 0.0f;
    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
    for (int t = 0; t < NT; ++t) {
        const int m = t * 128 + rowid;
        const int r = m / kRowsPerRoot, q = m - r * kRowsPerRoot, y = q / kPitch, x = q - y * kPitch;
        const bool valid = (m < rows_used) && (r < nvalid) && (y < 6) && (x < 6);
        const int p = y * 6 + x;
        float v[16];
        // warps 0-3 scatter reward + value features, warps 4-7 policy features (any warp may read its lane quarter)
        if (half == 0 && (hmask & 1)) {
            tmem_ld16(lane_base + kColRew + t * 16, v);
            if (valid)
                for (int c = 0; c < net.hc[0]; ++c)
                    hflat[(slot[0] * kMaxRoots + r) * 576 + c * kP + p] = fmaxf(fmaf(v[c], __ldg(net.head_bn + c), __ldg(net.head_bn + 16 + c)), 0.0f);
        }
        if (half == 0 && (hmask & 2)) {
            tmem_ld16(lane_base + kColAcc + t * 32, v);
            if (valid)
                for (int c = 0; c < net.hc[1]; ++c)
                    hflat[(slot[1] * kMaxRoots + r) * 576 + c * kP + p] = fmaxf(fmaf(v[c], __ldg(net.head_bn + 32 + c), __ldg(net.head_bn + 48 + c)), 0.0f);
        }
        if (half == 1 && (hmask & 4)) {
            tmem_ld16(lane_base + kColAcc + t * 32 + 16, v);
            if (valid)
                for (int c = 0; c < net.hc[2]; ++c)
                    hflat[(slot[2] * kMaxRoots + r) * 576 + c * kP + p] = fmaxf(fmaf(v[c], __ldg(net.head_bn + 64 + c), __ldg(net.head_bn + 80 + c)), 0.0f);
        }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
    if (hmask == 1 && io.ez_feat) {
        // EfficientZero (efficientzero_model.py:556-562): the flattened reward features feed an LSTM that is evaluated as
        // one batched GEMM over all roots by the next kernel (ez.cu)
        const int nin = net.hc[0] * kP;
        for (int i = tid; i < nvalid * nin; i += kEpiThreads) {
            const int r = i / nin, j = i - r * nin;
            io.ez_feat[(size_t)(root0 + r) * nin + j] = hflat[r * 576 + j];
        }
        return;
    }
    // ---- FC1 (hc*36 -> hid): warp w takes an eighth of the inputs, lane = hidden unit; 24 coalesced weight rows in flight
    for (int h = 0; h < 3; ++h) {
        if (!((hmask >> h) & 1)) continue;
        const Head &H = h == 0 ? net.reward : (h == 1 ? net.value : net.policy);
        const int nin = H.hc * kP, qn = (nin + kEpiWarps - 1) / kEpiWarps, i0 = warp * qn, i1 = min(nin, i0 + qn);
        float a[kMaxRoots];
#pragma unroll
        for (int r = 0; r < kMaxRoots; ++r) a[r] = 0.0f;
        const float *hf = hflat + slot[h] * kMaxRoots * 576;
        const float *wp = H.fc1 + lane;
        const bool lane_on = lane < H.hid;
        for (int i = i0; i < i1; i += 24) {
            float w[24];
#pragma unroll
            for (int u = 0; u < 24; ++u) w[u] = (lane_on && i + u < i1) ? __ldg(wp + (size_t)(i + u) * H.hid) : 0.0f;
#pragma unroll
            for (int u = 0; u < 24; ++u) {
                const int ii = min(i + u, nin - 1);
#pragma unroll
                for (int r = 0; r < kMaxRoots; ++r) a[r] = fmaf(hf[r * 576 + ii], w[u], a[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kMaxRoots; ++r) part[((warp * nh + slot[h]) * kMaxRoots + r) * 32 + lane] = a[r];
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
    for (int o = tid; o < nh * kMaxRoots * 32; o += kEpiThreads) {
        const int sl = o / (kMaxRoots * 32), j = o & 31;
        const int h = (sl == slot[0] && (hmask & 1)) ? 0 : ((sl == slot[1] && (hmask & 2)) ? 1 : 2);
        const Head &H = h == 0 ? net.reward : (h == 1 ? net.value : net.policy);
        float v = 0.0f;
#pragma unroll
        for (int w8 = 0; w8 < kEpiWarps; ++w8) v += part[o + w8 * nh * kMaxRoots * 32];
        hidden[o] = (j < H.hid) ? fmaxf(fmaf(v, __ldg(H.s2 + j), __ldg(H.t2 + j)), 0.0f) : 0.0f;
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
    // ---- FC2 (hid -> K): thread = output k, weights [j][K] coalesced over k, all 32 rows in flight
    for (int h = 0; h < 3; ++h) {
        if (!((hmask >> h) & 1)) continue;
        const Head &H = h == 0 ? net.reward : (h == 1 ? net.value : net.policy);
        float *lg = logits + slot[h] * kMaxRoots * ldl;
        const float *hid = hidden + slot[h] * kMaxRoots * 32;
        for (int k = tid; k < H.K; k += kEpiThreads) {
            float o[kMaxRoots];
            const float bias = __ldg(H.b2 + k);
#pragma unroll
            for (int r = 0; r < kMaxRoots; ++r) o[r] = bias;
            float w[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) w[u] = (u < H.hid) ? __ldg(H.fc2 + (size_t)u * H.K + k) : 0.0f;
#pragma unroll
            for (int u = 0; u < 32; ++u) {
#pragma unroll
                for (int r = 0; r < kMaxRoots; ++r) o[r] = fmaf(hid[r * 32 + u], w[u], o[r]);
            }
#pragma unroll
            for (int r = 0; r < kMaxRoots; ++r) lg[r * ldl + k] = o[r];
        }
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
    // ---- softmax expectation + inverse transform: one warp per (categorical head, root)
    for (int task = warp; task < 2 * kMaxRoots; task += kEpiWarps) {
        const int h = task / kMaxRoots, r = task - h * kMaxRoots;       // h: 0 reward, 1 value (+ policy copy)
        if (r >= nvalid || !((hmask >> h) & 1)) continue;
        const int b = root0 + r;
        const float *lg = logits + (slot[h] * kMaxRoots + r) * ldl;
        if (h == 0) {
            const float rv = categorical_to_scalar(lg, net.reward.K, net.support_min, net.support_step, lane);
            if (lane == 0 && io.reward) io.reward[b] = rv;
            if (io.reward_logits)
                for (int k = lane; k < net.reward.K; k += 32) io.reward_logits[(size_t)b * net.reward.K + k] = lg[k];
        } else {
            const float vv = categorical_to_scalar(lg, net.value.K, net.support_min, net.support_step, lane);
            if (lane == 0 && io.value) io.value[b] = vv;
            if (io.value_logits)
                for (int k = lane; k < net.value.K; k += 32) io.value_logits[(size_t)b * net.value.K + k] = lg[k];
            if (io.policy_logits && (hmask & 4)) {
                const float *lp = logits + (slot[2] * kMaxRoots + r) * ldl;
                for (int a = lane; a < A; a += 32) io.policy_logits[(size_t)b * A + a] = lp[a];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kTcThreads, 1) k_net_tc(TcNet net, TcIO io, TreeParams tp)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *act = smem;                                   // [2 parts][8 planes][400 rows][16 B]
    unsigned char *ring = smem + kActBytes;                      // [kStages][hi 8 KB | lo 8 KB]
    unsigned char *headw = ring + kStages * kTapBytes;           // [3 heads][hi 2 KB | lo 2 KB]
    TcBars *bars = reinterpret_cast<TcBars *>(headw + kHeadWBytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = io.roots_per_cta;
    const int root0 = blockIdx.x * R;
    const int nvalid = min(R, io.B - root0);                     // roots of this CTA that exist
    const int rows_used = R * kRowsPerRoot;
    const int NT = (rows_used + 127) >> 7;
    const int npass = io.npass;
    const int nlayers = net.nlayers;
    const int nsims = io.nsims > 0 ? io.nsims : 1;     // > 1 (or persistent): the whole search loop runs inside this launch
    const bool persistent = io.persistent != 0;

    pdl_launch_dependents();      // the next tree kernel may become resident; it blocks in pdl_wait() until this grid is done
    // ---- one-time setup ----
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        mbar_init(&bars->acc_ready, 1);
        mbar_init(&bars->act_ready, kEpiThreads);
        mbar_init(&bars->rew_ready, 1);
        mbar_init(&bars->vp_ready, 1);
        fence_mbar_init();
    }
    if (warp == kEpiWarps + 1) tmem_alloc(&bars->tmem_base, kTmemCols);
    // 1x1 head weights: plain copy (12 KB)
    for (int i = tid; i < kHeadWBytes / 16; i += kTcThreads)
        reinterpret_cast<uint4 *>(headw)[i] = __ldg(reinterpret_cast<const uint4 *>(net.headw) + i);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == kEpiWarps) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t n = 0;
            for (int sim = 0; sim < nsims; ++sim)       // runs ahead of the MMAs: the next simulation's first taps are
                for (int L = 0; L < nlayers; ++L) {     // already in the ring while heads / tree work is going on
                    const unsigned char *src = net.convw + (size_t)net.layer_w[L] * (9 * kTapBytes);
                    for (int tap = 0; tap < 9; ++tap, ++n) {
                        const int st = n % kStages;
                        if (n >= kStages) mbar_wait(&bars->empty[st], ((n / kStages) - 1) & 1);
                        mbar_expect_tx(&bars->full[st], kTapBytes);
                        bulk_g2s(ring + st * kTapBytes, src + (size_t)tap * kTapBytes, kTapBytes, &bars->full[st]);
                    }
                }
        }
    } else if (warp == kEpiWarps + 1) {
        // ================= MMA issuer =================
        if (LZ_MMA_ISSUER_ON) {
            const uint32_t act_s = smem_u32(act), ring_s = smem_u32(ring), headw_s = smem_u32(headw);
            const uint32_t idesc64 = make_idesc_f16(128, 64), idesc16 = make_idesc_f16(128, 16), idesc32 = make_idesc_f16(128, 32);
            const bool sw = false;
            const uint32_t a_lbo = kPlaneBytes >> 4, a_sbo = 8;
            const uint64_t a_desc0 = make_desc(act_s + kMargin * 16, a_lbo, a_sbo);   // row 0, hi part, k-step 0
            const uint64_t b_desc0 = make_desc(ring_s, 64, 8);                        // stage 0, hi part, k-step 0
            unsigned long long *dbg = (io.dbg && blockIdx.x == 0) ? io.dbg : nullptr;
            uint32_t n = 0;
            for (int sim = 0; sim < nsims; ++sim)
            for (int L = 0; L < nlayers; ++L) {
                const uint32_t ev = (uint32_t)sim * (nlayers + 1) + L;   // act_ready event index: 1 load + nlayers epilogues per simulation
                mbar_wait(&bars->act_ready, ev & 1);             // inputs written, TMEM accumulators drained
                tc_fence_after();
                if (dbg && sim == 0) dbg[32 + 2 * L] = clock64();
                for (int tap = 0; tap < 9; ++tap, ++n) {
                    const int st = n % kStages;
                    mbar_wait(&bars->full[st], (n / kStages) & 1);
                    tc_fence_after();
                    const int shift = (tap / 3 - 1) * kPitch + (tap % 3 - 1);
                    // descriptors differ only in the 14-bit start-address field: add 16-byte-unit offsets to a base
                    const uint64_t b0 = b_desc0 + (uint64_t)((st * kTapBytes) >> 4);
                    for (int t = 0; t < NT; ++t) {
                        const uint64_t a0 = a_desc0 + (uint64_t)(t * 128 + shift);
                        const uint32_t d = tmem + kColAcc + t * 64;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            LZ_UMMA(d, a0 + ks * (2 * kPlaneBytes >> 4), b0 + ks * (2048 >> 4), idesc64, (tap | ks) != 0);
                        if (npass == 3) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)      // A_hi * B_lo
                                LZ_UMMA(d, a0 + ks * (2 * kPlaneBytes >> 4), b0 + ((kTapBytes / 2) >> 4) + ks * (2048 >> 4), idesc64, 1);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)      // A_lo * B_hi
                                LZ_UMMA(d, a0 + (kPartBytes >> 4) + ks * (2 * kPlaneBytes >> 4), b0 + ks * (2048 >> 4), idesc64, 1);
                        }
                    }
                    LZ_UCOMMIT(&bars->empty[st]);               // frees this ring slot when the MMAs have read it
                }
                LZ_UCOMMIT(&bars->acc_ready);
                if (dbg && sim == 0) dbg[33 + 2 * L] = clock64();
                const int flags = net.layer_flags[L];
                if (flags & (LF_HOOK_REWARD | LF_HOOK_VALPOL)) {
                    // 1x1 head convolutions on this layer's OUTPUT: wait for the epilogue to have written it (the same
                    // phase the next layer waits for; waiting twice on a completed phase is immediate).  The
                    // value/policy result reuses the drained conv accumulator columns, so that hook is only legal on
                    // the LAST layer; the reward result has its own columns.
                    mbar_wait(&bars->act_ready, (ev + 1) & 1);
                    tc_fence_after();
                    for (int hook = 0; hook < 2; ++hook) {
                        if (!(flags & (hook == 0 ? LF_HOOK_REWARD : LF_HOOK_VALPOL))) continue;
                        for (int t = 0; t < NT; ++t) {
                            const uint32_t arow = act_s + (uint32_t)(kMargin + t * 128) * 16u;
                            for (int ps = 0; ps < npass; ++ps) {
                                const uint32_t apart = (ps == 2) ? kPartBytes : 0;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint64_t ad = make_desc(arow + apart + ks * 2 * kPlaneBytes, a_lbo, a_sbo);
                                    if (hook == 0) {   // reward head: N = 16, [kg][16 co][8]
                                        const uint32_t wb = headw_s + ((ps == 1) ? 2048 : 0);
                                        LZ_UMMA(tmem + kColRew + t * 16, ad, make_desc(wb + ks * 512, sw ? 8 : 16, sw ? 16 : 8), idesc16, (ps | ks) != 0);
                                    } else {           // value + policy heads as one N = 32 matrix, [kg][32 co][8]
                                        const uint32_t wb = headw_s + 4096 + ((ps == 1) ? 4096 : 0);
                                        LZ_UMMA(tmem + kColAcc + t * 32, ad, make_desc(wb + ks * 1024, sw ? 8 : 32, sw ? 32 : 8), idesc32, (ps | ks) != 0);
                                    }
                                }
                            }
                        }
                        LZ_UCOMMIT(hook == 0 ? &bars->rew_ready : &bars->vp_ready);
                    }
                }
            }
        }
    } else {
        // ================= epilogue warps: warp w owns TMEM lanes 32*(w%4).. and the 32-column half w/4 =================
        const int q4 = warp & 3, half = warp >> 2, rowid = q4 * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
        unsigned long long *dbg = (io.dbg && blockIdx.x == 0 && tid == 0) ? io.dbg : nullptr;
        if (dbg) dbg[0] = clock64();
        pdl_wait();                   // ix / action / the latent pool come from the preceding kernels
        int acc_par = 0;
        for (int sim = 0; sim < nsims; ++sim) {
            if (persistent) {
                // ---- tree phase: one warp per root of this CTA (roots never interact, so the whole search of these roots
                // lives in this CTA): back up the previous simulation, then descend to the next leaf.  cnode.cpp:480-500,754-825
                if (warp < nvalid) {
                    const int b = root0 + warp;
                    if (sim > 0)
                        tree_backprop(tp, b, lane, io.sim0 + sim, io.reward[b], io.value[b], io.policy_logits + (size_t)b * net.A, nullptr);
                    tree_traverse(tp, b, lane, io.deterministic, (unsigned)(io.sim0 + sim), io.ix_rw, nullptr, io.action_rw, nullptr, nullptr);
                }
                __threadfence_block();
                asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
            }
            float *latent_out = persistent ? (io.latent_pool_rw + (size_t)(io.sim0 + sim + 1) * io.slot_stride) : io.latent_out;
            // zero the margins (the head scratch of the previous simulation overlays them; pad rows inside the tiles are
            // rewritten as zeros by every load / epilogue)
            for (int i = tid; i < 2 * 8 * 2 * kMargin; i += kEpiThreads) {
                int part = i / (8 * 2 * kMargin), rem = i % (8 * 2 * kMargin), plane = rem / (2 * kMargin), r = rem % (2 * kMargin);
                int row = r < kMargin ? r : kMargin + kMaxTiles * 128 + (r - kMargin);
                *reinterpret_cast<uint4 *>(act + part * kPartBytes + plane * kPlaneBytes + row * 16) = make_uint4(0, 0, 0, 0);
            }
            // ---- load the input activation: gather NCHW latents, split to fp16 hi/lo, park fp32 copy in TMEM ----
            for (int t = 0; t < NT; ++t) {
                const int m = t * 128 + rowid;
                const int r = m / kRowsPerRoot, q = m - r * kRowsPerRoot, y = q / kPitch, x = q - y * kPitch;
                const bool valid = (m < rows_used) && (r < nvalid) && (y < 6) && (x < 6);
                const float *src = nullptr;
                if (valid) {
                    const int b = root0 + r;
                    const size_t slot = io.ix ? (size_t)io.ix[b] : 0;
                    src = io.latent_base + slot * io.slot_stride + (size_t)b * (kC * kP) + (y * 6 + x);
                }
                float v[32];
    #pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = valid ? src[(size_t)(half * 32 + c) * kP] : 0.0f   /* plain load: the pool is written by this launch in persistent mode */;
    #pragma unroll
                for (int g = 0; g < 4; ++g) {
                    unsigned char *p = act + (half * 4 + g) * kPlaneBytes + (kMargin + m) * 16;
                    store_split8(p, p + kPartBytes, v + 8 * g);
                }
                tmem_st32(lane_base + kColRes + t * 64 + half * 32, v);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->act_ready);                          // phase 0: layer 0 may start
            if (dbg) dbg[1] = clock64();


            for (int L = 0; L < nlayers; ++L) {
                const int flags = net.layer_flags[L];
                const float *bn = net.bn + (size_t)net.layer_w[L] * 128;      // [scale 64 | shift 64]
                mbar_wait_warp(&bars->acc_ready, acc_par);
                acc_par ^= 1;                                       // one commit per layer, across simulations
                tc_fence_after();
                if (dbg) dbg[2 + 2 * L] = clock64();
                float bs[32], bt[32];                               // folded BatchNorm of this thread's 32 channels
    #pragma unroll
                for (int c = 0; c < 32; ++c) { bs[c] = __ldg(bn + half * 32 + c); bt[c] = __ldg(bn + 64 + half * 32 + c); }
                for (int t = 0; t < NT; ++t) {
                    const int m = t * 128 + rowid;
                    const int r = m / kRowsPerRoot, q = m - r * kRowsPerRoot, y = q / kPitch, x = q - y * kPitch;
                    const bool valid = (m < rows_used) && (r < nvalid) && (y < 6) && (x < 6);
                    const int b = root0 + r, p = y * 6 + x;
                    float v[32];
                    tmem_ld32(lane_base + kColAcc + t * 64 + half * 32, v);
                    if (flags & LF_RES) {
                        float rs[32];
                        tmem_ld32(lane_base + kColRes + t * 64 + half * 32, rs);
    #pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = fmaf(v[c], bs[c], bt[c]) + rs[c];
                    } else {
    #pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = fmaf(v[c], bs[c], bt[c]);
                    }
                    if ((flags & LF_ACT_BIAS) && valid) {
                        const int action = min(max(io.action[b], 0), net.A - 1);
                        const float *ab = net.abias + ((size_t)action * kC + half * 32) * kP + p;
    #pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] += __ldg(ab + (size_t)c * kP);
                    }
    #pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = valid ? fmaxf(v[c], 0.0f) : 0.0f;
                    if (flags & LF_STORE_RES) tmem_st32(lane_base + kColRes + t * 64 + half * 32, v);
                    if ((flags & LF_WRITE_LATENT) && valid) {
                        if (latent_out) {
                            float *dst = latent_out + (size_t)b * (kC * kP) + (size_t)(half * 32) * kP + p;
    #pragma unroll
                            for (int c = 0; c < 32; ++c) dst[(size_t)c * kP] = v[c];
                        }
                        if (io.latent_out2) {
                            float *dst = io.latent_out2 + (size_t)b * (kC * kP) + (size_t)(half * 32) * kP + p;
    #pragma unroll
                            for (int c = 0; c < 32; ++c) dst[(size_t)c * kP] = v[c];
                        }
                    }
    #pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        unsigned char *pp = act + (half * 4 + g) * kPlaneBytes + (kMargin + m) * 16;
                        store_split8(pp, pp + kPartBytes, v + 8 * g);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&bars->act_ready);                      // phase L+1: next layer / this layer's hook may start
                if (dbg) dbg[3 + 2 * L] = clock64();
                if ((flags & LF_HOOK_REWARD) && net.has_reward_early) {
                    // the reward head (FC1, FC2, softmax, h^-1) runs here, underneath the NEXT layer's MMAs
                    mbar_wait_warp(&bars->rew_ready, sim & 1);
                    tc_fence_after();
                    head_stage(net, io, 1, reinterpret_cast<float *>(smem + kSmemMain), tmem, NT, rows_used, nvalid, root0);
                    if (dbg) dbg[28] = clock64();
                }
            }

            // all 1x1 head accumulators must be complete before the head stage reads them / reuses the buffer
            if (net.has_reward && !net.has_reward_early) mbar_wait_warp(&bars->rew_ready, sim & 1);
            mbar_wait_warp(&bars->vp_ready, sim & 1);
            tc_fence_after();
            if (dbg) dbg[24] = clock64();
            // ---- heads: 1x1 accumulators -> BN/ReLU -> FC1 -> FC2 -> softmax expectation -> h^-1 (scratch overlays the
            // activation buffer: every conv MMA of this simulation has completed)
            head_stage(net, io, net.has_reward_early ? 6 : (net.has_reward ? 7 : 6), reinterpret_cast<float *>(act), tmem, NT, rows_used, nvalid, root0);
            if (dbg) dbg[27] = clock64();
            dbg = nullptr;
            __threadfence_block();
            asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
        }
        if (persistent && warp < nvalid) {        // back up the last simulation (mcts_ctree.py:365-368)
            const int b = root0 + warp;
            tree_backprop(tp, b, lane, io.sim0 + nsims, io.reward[b], io.value[b], io.policy_logits + (size_t)b * net.A, nullptr);
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

// One 3x3 conv -> 9 tap blocks of [hi | lo], each [kg = ci/8][co 64][ci % 8] fp16.  Returns the power-of-two
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
                const size_t off = (size_t)t * (kTapBytes / 2) + ((size_t)(ci / 8) * 64 + co) * 8 + (ci % 8);
                h[off] = hi;
                h[off + kTapBytes / 4] = lo;      // lo block follows the 8 KB hi block (4096 halves)
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

int tc_prepare_launch()
{
    LZ_CUDA_CHECK(cudaFuncSetAttribute(k_net_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return LZ_OK;
}

int tc_pick_roots(int B)
{
    int r = (B + 147) / 148;
    return std::min(std::max(r, 1), kMaxRoots);
}

static unsigned long long *g_dbg = nullptr;
unsigned long long *tc_debug_buffer() { return g_dbg; }

int tc_launch(const TcNet &net, const TcIO &io_in, cudaStream_t s, const TreeParams *tp_in)
{
    TcIO io = io_in;
    TreeParams tp;
    memset(&tp, 0, sizeof(tp));
    if (tp_in) tp = *tp_in;
    LZ_REQUIRE(!io.persistent || (tp_in && net.has_reward_early), LZ_EINVAL, "tc_launch: persistent search needs tree parameters and the early reward head");
    if (getenv("LZ_TC_DEBUG")) {
        if (!g_dbg) { cudaMalloc(&g_dbg, 64 * 8); cudaMemset(g_dbg, 0, 64 * 8); }
        io.dbg = g_dbg;
    }
    io.roots_per_cta = tc_pick_roots(io.B);
    if (const char *e = getenv("LZ_TC_VARIANT")) io.variant = atoi(e);
    if (const char *e = getenv("LZ_TC_ROOTS")) io.roots_per_cta = std::min(std::max(atoi(e), 1), kMaxRoots);
    const int grid = (io.B + io.roots_per_cta - 1) / io.roots_per_cta;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = io_in.pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    LZ_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_net_tc, net, io, tp));
    return LZ_OK;
}

}  // namespace lz
