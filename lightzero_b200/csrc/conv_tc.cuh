// conv_tc.cuh -- tcgen05 3x3 convolutions of the DownSample tower (representation network).
//
// Activations between tower layers live in HBM in a tensor-core-ready layout ("TCL"): per image
//   [part: fp16 hi | fp16 lo][k-group of 8 channels][row][8 halves]
// where `row` walks a zero-padded grid of pitch W+1 (one pad column) with one pad row above and below:
//   rho = (y + 1) * pitch + x,   memory row = rho + 1   (one spare zero row at each end),
//   rows per plane R = (H + 2) * pitch + 2.
// A band of image rows is then a handful of contiguous cp.async.bulk copies per plane, lands in shared
// memory already in the UMMA K-major / no-swizzle canonical layout, and every 3x3 tap is the same buffer
// read through a row-shifted descriptor (see net_tc.cu).  Stride-2 convolutions read a 4-phase
// (space-to-depth) variant written by the producing layer: phase (y&1, x&1) image of half size, so that
// tap (ky,kx) is phase ((ky+1)&1, (kx+1)&1) shifted by (ky==0 ? -1 : 0, kx==0 ? -1 : 0).
#pragma once
#include "lz_common.cuh"

namespace lz {

struct Tcl {                      // one activation tensor in TCL
    unsigned char *base;
    size_t img_stride, phase_stride, part_stride;   // bytes
    int plane_rows;               // R (rows per k-group plane, incl. the 2 spare rows)
    int C, H, W, pitch;           // geometry of ONE phase image (== the tensor itself when nphase == 1)
    int nphase;                   // 1 or 4
};

inline Tcl make_tcl(unsigned char *base, int C, int H, int W, int nphase)
{
    Tcl t;
    t.base = base; t.C = C; t.H = H; t.W = W; t.pitch = W + 1; t.nphase = nphase;
    t.plane_rows = (H + 2) * (W + 1) + 2;
    t.part_stride = (size_t)(C / 8) * t.plane_rows * 16;
    t.phase_stride = 2 * t.part_stride;
    t.img_stride = t.phase_stride * nphase;
    return t;
}
inline size_t tcl_bytes(int B, int C, int H, int W, int nphase)
{
    return (size_t)B * nphase * 2 * (C / 8) * ((size_t)(H + 2) * (W + 1) + 2) * 16;
}

struct ConvTc {
    Tcl in;                       // input (nphase 1 or 4); its (H, W, pitch) is the output/row-space geometry
    Tcl out[2];                   // column group 0 / 1 (group 1 only for N = 128); nphase 4 = write phase-split
    Tcl res;                      // residual added to group 0 (base == nullptr: none)
    const unsigned char *w;       // [9 taps][kg_in][N hi rows | N lo rows][8] fp16
    const float *scale, *shift;   // [N]
    int tap_phase[9], tap_shift[9];
    int relu[2];
    int N;                        // 32, 64 or 128
    int G, band_h;                // images per CTA (band_h == H when G > 1), image rows per band
    int stages;                   // weight ring depth (2..4)
    int fold;                     // fp32-accurate mode: A_hi x [B_hi | B_lo] as ONE 2N-column MMA (N <= 64; 2N accumulator columns per tile)
    int B, npass;
    unsigned long long *dbg;      // bring-up instrumentation (env LZ_CONV_DEBUG=<layer>): clock64 stamps of the middle CTA, slots 58-63 of the debug buffer
};

int conv_tc_prepare_launch();
int conv_tc_launch(const ConvTc &p, cudaStream_t s);
// TMEM accumulator columns per 128-row tile: layers with N <= 64 output columns fold [B_hi | B_lo] into one 2N-column MMA (conv_tc.cu)
inline int conv_tc_acc_cols(const ConvTc &p) { return (p.fold && p.N <= 64) ? 2 * p.N : p.N; }
// weights [cout][cin][3][3] -> tap blocks with `ncols` columns, this tensor occupying columns
// [col0, col0+cout); returns the exact power-of-two scale applied
float conv_tc_pack(const float *w_torch, int cin, int cout, int ncols, int col0, unsigned char *dst);
size_t conv_tc_packed_bytes(int cin, int ncols);

// pooling on TCL tensors: AvgPool2d(3, stride 2, pad 1), count_include_pad -> /9
int pool_tcl_launch(const Tcl &in, const Tcl &out, int B, cudaStream_t s);                 // TCL -> TCL
int pool_tcl_to_nchw_launch(const Tcl &in, float *out, int B, int Hout, cudaStream_t s);   // TCL -> fp32 NCHW

}  // namespace lz
