// ez.cu -- EfficientZero value-prefix head for sm_100a: LSTM step + BatchNorm1d/ReLU/MLP/categorical expectation.
//
// k_ez_lstm: gates[B][4H] = [feat | h_in] (B x (nin+H)) * wcat ((nin+H) x 4H) + bias as a tiled fp32 GEMM over ALL roots
// (the weights, 8.9 MB at nin = 576 / H = 512, are read once per 64-row tile instead of once per root), with the LSTM cell
// update fused into the epilogue: the weight columns are ordered unit-major / gate-minor so that the 4 x 4 register tile
// of a thread holds (i, f, g, o) of one hidden unit for 4 roots.  torch.nn.LSTM gate order i, f, g, o; c' = sig(f) c +
// sig(i) tanh(g); h' = sig(o) tanh(c').
// k_ez_head: per root relu(bn(h')) -> Linear(H, hid) + BN + ReLU -> Linear(hid, K) -> softmax expectation -> h^-1.
#include "ez.cuh"
#include "lz_common.cuh"

namespace lz {

constexpr int kGM = 64, kGN = 32, kGK = 16;      // 64 roots x 32 gate columns (= 8 hidden units) per CTA: 64 x ceil(B/64) CTAs
constexpr int kGThreads = 128;                   // thread = 4 roots x 4 gates of one hidden unit

__global__ void __launch_bounds__(kGThreads) k_ez_lstm(EzNet net, EzIO io)
{
    __shared__ float As[kGK][kGM + 4];
    __shared__ float Bs[kGK][kGN];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;          // tx: hidden unit within the tile, ty: 4-row group
    const int n0 = blockIdx.x * kGN, m0 = blockIdx.y * kGM;
    const int H = net.H, nin = net.nin, KT = nin + H, N = 4 * H;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    // A-tile loader: thread -> (row = tid / 2, 8 consecutive k); B-tile loader: thread -> (k = tid / 8, 4 consecutive n)
    const int ar = tid >> 1, ak = (tid & 1) * 8;
    const int arow = m0 + ar;
    const bool arow_on = arow < io.B;
    const float *hsrc = io.h_base + (arow_on && io.ix ? (size_t)io.ix[arow] * io.slot_stride : 0) + (size_t)(arow_on ? arow : 0) * H;
    const float *fsrc = io.feat + (size_t)(arow_on ? arow : 0) * nin;
    const int bk = tid >> 3, bn = (tid & 7) * 4;
    float4 a4[2];
    float4 b4;
    auto fetch = [&](int k0) {       // global -> registers (in flight while the previous tile is being multiplied)
#pragma unroll
        for (int v = 0; v < 2; ++v) {        // nin and H are multiples of 4, so a 16-byte load never straddles feat | h
            const int k = k0 + ak + 4 * v;
            a4[v] = (arow_on && k < KT) ? *reinterpret_cast<const float4 *>(k < nin ? fsrc + k : hsrc + (k - nin)) : make_float4(0, 0, 0, 0);
        }
        b4 = (k0 + bk < KT) ? __ldg(reinterpret_cast<const float4 *>(net.wcat + (size_t)(k0 + bk) * N + n0 + bn)) : make_float4(0, 0, 0, 0);
    };
    fetch(0);
    for (int k0 = 0; k0 < KT; k0 += kGK) {
        __syncthreads();
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            As[ak + 4 * v + 0][ar] = a4[v].x; As[ak + 4 * v + 1][ar] = a4[v].y;
            As[ak + 4 * v + 2][ar] = a4[v].z; As[ak + 4 * v + 3][ar] = a4[v].w;
        }
        *reinterpret_cast<float4 *>(&Bs[bk][bn]) = b4;
        __syncthreads();
        if (k0 + kGK < KT) fetch(k0 + kGK);
#pragma unroll
        for (int k = 0; k < kGK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    // epilogue: this thread owns hidden unit `unit` for rows m0 + ty*4 .. +3
    const int unit = (n0 >> 2) + tx;
    const float4 bias = *reinterpret_cast<const float4 *>(net.bias + n0 + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = m0 + ty * 4 + i;
        if (b >= io.B) continue;
        const float gi = acc[i][0] + bias.x, gf = acc[i][1] + bias.y, gg = acc[i][2] + bias.z, go = acc[i][3] + bias.w;
        const float c_in = io.c_base[(io.ix ? (size_t)io.ix[b] * io.slot_stride : 0) + (size_t)b * H + unit];
        const float si = 1.0f / (1.0f + expf(-gi)), sf = 1.0f / (1.0f + expf(-gf)), so = 1.0f / (1.0f + expf(-go));
        const float c_new = sf * c_in + si * tanhf(gg);
        const float h_new = so * tanhf(c_new);
        const bool reset = io.is_reset && io.is_reset[b] != 0;
        io.h_tmp[(size_t)b * H + unit] = h_new;
        if (io.h_out) io.h_out[(size_t)b * H + unit] = reset ? 0.0f : h_new;
        if (io.c_out) io.c_out[(size_t)b * H + unit] = reset ? 0.0f : c_new;
    }
}

constexpr int kHR = 2;        // roots per CTA of the head kernel (the FC weights are 64 KB + 77 KB, L2-resident)
constexpr int kHMaxH = 512, kHMaxHid = 32, kHLd = 608;

__global__ void __launch_bounds__(256) k_ez_head(EzNet net, EzIO io)
{
    __shared__ float x[kHR][kHMaxH];
    __shared__ float part[8][kHR][kHMaxHid];
    __shared__ float hidden[kHR][kHMaxHid];
    __shared__ float logits[kHR][kHLd];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * kHR, nr = min(kHR, io.B - r0);
    const int H = net.H, hid = net.hid, K = net.K;
    for (int i = tid; i < kHR * H; i += 256) {                       // norm_value_prefix + ReLU (efficientzero_model.py:565-566)
        const int r = i / H, u = i - r * H;
        x[r][u] = r < nr ? fmaxf(fmaf(io.h_tmp[(size_t)(r0 + r) * H + u], net.vp_s[u], net.vp_t[u]), 0.0f) : 0.0f;
    }
    __syncthreads();
    {   // Linear(H -> hid): warp w sums its eighth of the inputs, lane = hidden unit; 16 weight rows in flight per batch
        const int per = (H + 7) / 8, i0 = warp * per, i1 = min(H, i0 + per);
        float a[kHR];
#pragma unroll
        for (int r = 0; r < kHR; ++r) a[r] = 0.0f;
        const bool lane_on = lane < hid;
        for (int i = i0; i < i1; i += 16) {
            float w[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) w[u] = (lane_on && i + u < i1) ? __ldg(net.fc1 + (size_t)(i + u) * hid + lane) : 0.0f;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int ii = min(i + u, H - 1);
#pragma unroll
                for (int r = 0; r < kHR; ++r) a[r] = fmaf(x[r][ii], w[u], a[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kHR; ++r) part[warp][r][lane] = a[r];
    }
    __syncthreads();
    if (tid < kHR * 32) {
        const int r = tid >> 5, j = tid & 31;      // kHR roots x 32 hidden units
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += part[w][r][j];
        hidden[r][j] = j < hid ? fmaxf(fmaf(v, net.s2[j], net.t2[j]), 0.0f) : 0.0f;
    }
    __syncthreads();
    for (int k = tid; k < K; k += 256) {           // Linear(hid -> K)
        float o[kHR];
        const float bias = __ldg(net.b2 + k);
#pragma unroll
        for (int r = 0; r < kHR; ++r) o[r] = bias;
        float w[kHMaxHid];
#pragma unroll
        for (int j = 0; j < kHMaxHid; ++j) w[j] = j < hid ? __ldg(net.fc2 + (size_t)j * K + k) : 0.0f;     // all rows in flight
#pragma unroll
        for (int j = 0; j < kHMaxHid; ++j) {
#pragma unroll
            for (int r = 0; r < kHR; ++r) o[r] = fmaf(hidden[r][j], w[j], o[r]);
        }
#pragma unroll
        for (int r = 0; r < kHR; ++r) logits[r][k] = o[r];
    }
    __syncthreads();
    if (warp < nr) {
        const int b = r0 + warp;
        const float vp = categorical_to_scalar(logits[warp], K, net.support_min, net.support_step, lane);
        if (lane == 0 && io.value_prefix) io.value_prefix[b] = vp;
        if (io.vp_logits)
            for (int k = lane; k < K; k += 32) io.vp_logits[(size_t)b * K + k] = logits[warp][k];
    }
}

int ez_launch(const EzNet &net, const EzIO &io, cudaStream_t s)
{
    LZ_REQUIRE(net.H <= kHMaxH && net.hid <= kHMaxHid && net.K <= kHLd && (net.H % 8) == 0, LZ_EINVAL,
               "ez_launch: unsupported LSTM / head size (H=%d hid=%d K=%d)", net.H, net.hid, net.K);
    dim3 grid(4 * net.H / kGN, (io.B + kGM - 1) / kGM);
    k_ez_lstm<<<grid, kGThreads, 0, s>>>(net, io);
    LZ_KERNEL_CHECK();
    k_ez_head<<<(io.B + kHR - 1) / kHR, 256, 0, s>>>(net, io);
    LZ_KERNEL_CHECK();
    return LZ_OK;
}

}  // namespace lz
