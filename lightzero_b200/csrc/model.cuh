// model.cuh -- device-side weight tables of the MuZero conv model and the internal launch API.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "lz_common.cuh"
#include "net6.cuh"
#include "net_tc.cuh"
#include "conv_tc.cuh"
#include "ez.cuh"

namespace lz {

constexpr int kMaxResBlocks = 4;

struct NetDev {                       // passed by value to kernels
    Conv3 dyn_conv;                   // dynamics_network.conv (+ norm_common)
    Conv3 dyn_res[2 * kMaxResBlocks]; // dynamics_network.resblocks
    Conv3 pred_res[2 * kMaxResBlocks];
    Conv3 rep_res[2 * kMaxResBlocks]; // representation_network.resblocks (on the latent grid)
    Head reward, value, policy;
    int nres;
    int A;
    float support_min, support_step;
};

struct ConvG {                        // generic 3x3 conv of the DownSample tower
    const float *w;                   // [cin][9][cout]
    const float *scale, *shift;       // [cout]
    int cin, cout, stride, hin, win, hout, wout;
};

struct Dense {                        // y = act(scale * (W x) + shift); wt is input-major [in][out]
    const float *wt, *scale, *shift;
    int in, out, act, nact;           // act: 0 none, 1 ReLU, 2 GELU(tanh); nact: trailing one-hot action inputs
};

struct MlpNet {                       // MuZeroModelMLP (muzero_model_mlp.py)
    Dense e0, e1;                     // representation: Linear+BN+GELU, Linear (+ LayerNorm)
    Dense d1a, d1b, d2a, d2b;         // dynamics: fc_dynamics_1 (or fc_dynamics), fc_dynamics_2
    Dense r0, r1, pc0, pc1, v0, v1, p0, p1;
    const float *ln_w, *ln_b;
    int latent, obs_dim, A, res;
    float support_min, support_step;
};

struct RecIO {
    int B;
    const float *latent_base;         // latent source: base + ix[b]*slot_stride + b*C*P  (ix == nullptr: slot 0)
    const int *ix;
    size_t slot_stride;
    const int *action;                // [B]
    float *next_latent;               // [B][C][P] destination (pool slot or API buffer) or nullptr
    float *reward, *value;            // [B] scalars or nullptr
    float *policy_logits;             // [B][A] or nullptr
    float *reward_logits, *value_logits;   // [B][K] or nullptr
    int pdl;                          // programmatic dependent launch (search graph)
    float *skip_scratch;              // [B][2304] scratch of the tcgen05 path (nullptr: the model's own, lz_model::tc_skip)
    // EfficientZero (reward == value prefix): LSTM state in / out, see ez.cuh
    const float *h_base, *c_base;     // base + ix[b]*hslot_stride + b*H
    size_t hslot_stride;
    float *h_out, *c_out;
    const int *is_reset;
};

struct TailIO {
    int B;
    const float *pre_latent;          // [B][C][P] output of the DownSample tower
    float *latent;                    // [B][C][P] (NCHW) or nullptr
    float *latent2;                   // second copy (latent pool slot 0) or nullptr
    float *value;                     // [B] scalar or nullptr
    float *policy_logits;             // [B][A] or nullptr
    float *value_logits;              // [B][K] or nullptr
};

}  // namespace lz

struct lz_model {
    int kind;                         // 0 = conv MuZeroModel / EfficientZeroModel (cfg.efficientzero), 1 = MuZeroModelMLP
    lz::EzNet ez;                     // EfficientZero value-prefix head tables (device pointers into d_weights)
    float *ez_feat, *ez_htmp;         // [ws_B][hc*36], [ws_B][H] scratch between the conv kernel and the LSTM kernels
    int ez_B;
    unsigned char *d_ez_wtc;          // LSTM weights in the tcgen05 layout (ez.cu)
    int latent_floats;                // floats per root latent (64*36 or latent_dim)
    lz_mlp_config mcfg;
    lz::MlpNet mlp;
    lz_model_config cfg;
    std::map<std::string, std::vector<float>> tensors;   // raw reference state_dict (host)
    bool finalized;
    float *d_weights;                 // one packed device allocation
    size_t n_weight_floats;
    lz::NetDev net;
    std::vector<lz::ConvG> tower;     // DownSample convs in execution order
    std::vector<float> stem_params;   // host copy of the Cin = 4 stem's weights + folded BN (kernel-parameter operands of k_stem4_tcl, model.cu)
    int stem_valid;
    int hw, P, K;
    int math;                         // 0 = fp32 FFMA (net6.cuh), 1 = tcgen05 3xFP16 (fp32-accurate), 2 = tcgen05 fp16 single pass
    unsigned char *d_tc;              // packed fp16 hi/lo weights + tables of the tcgen05 path
    lz::TcNet tc_rec, tc_tail;
    float *tc_skip;                   // [tc_skip_B][2304] ResBlock skip scratch of k_net_tc for launches outside a search (model_reserve)
    int tc_skip_B;
    // tcgen05 DownSample tower: packed weights / folded BN per layer, TCL activation workspace
    unsigned char *d_tower;           // weights + scale/shift tables
    lz::ConvTc tower_tc[7];           // rb1.c1, rb1.c2, ds(c1+c3), ds.c2, rb2.c1, rb2.c2, rb3.c1 / rb3.c2 share [6]: see model.cu
    lz::ConvTc tower_tc_rb3[2];
    unsigned char *tws;               // TCL workspace (one allocation)
    size_t tws_bytes;
    lz::Tcl T0, T1, T2, U0, U1, U2, V0, V1, V2;
    // workspace for initial inference (grown on demand, outside graph capture)
    float *ws[3];
    size_t ws_floats;
    int ws_B;
    unsigned long long generation;    // bumped when device tables / workspaces that captured search graphs point into are re-allocated
                                      // or the math mode changes (finalize, set_math, model_reserve): lz_search re-captures
};

namespace lz {
int model_recurrent(lz_model *m, const RecIO &io, cudaStream_t s);
int model_initial(lz_model *m, int B, const float *d_obs, const TailIO &io, cudaStream_t s);
int model_initial_tower(lz_model *m, int B, const float *d_obs, float *pre_latent, cudaStream_t s, const uint8_t *d_obs_u8 = nullptr);   // exactly one of d_obs / d_obs_u8
int model_initial_tail(lz_model *m, int B, const float *pre_latent, const TailIO &io, cudaStream_t s);
int model_reserve(lz_model *m, int B);   // sizes the initial-inference workspace (synchronous)
int mlp_recurrent(lz_model *m, const RecIO &io, cudaStream_t s);
int mlp_initial(lz_model *m, int B, const float *d_obs, const TailIO &io, cudaStream_t s);
int mlp_finalize(lz_model *m);
}  // namespace lz
