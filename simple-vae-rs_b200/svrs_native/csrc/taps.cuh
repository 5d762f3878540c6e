// taps.cuh - the "multi-tap GEMM" problem description shared by every convolution form.
//
// Every conv on the path (k3s1p1 fprop/dgrad, k4s2p1 fprop/dgrad, transposed k4s2p1 fprop/dgrad) is
//     out_view[n, oy, ox, :] = sum_t  in_view_t[n, oy + dy_t, ox + dx_t, :] @ W_t
// over a list of taps t, where out_view / in_view_t are strided views (full tensor, or one of the four
// stride-2 parity classes of a tensor) and reads outside [0,IH)x[0,IW) contribute zero.
//   conv k3 s1 p1 fprop : 1 problem, 9 taps, dy = ky-1, dx = kx-1
//   conv k3 s1 p1 dgrad : 1 problem, 9 taps, dy = 1-ky, dx = 1-kx (weights transposed by the pack)
//   conv k4 s2 p1 fprop : 1 problem, 16 taps; input row 2*oy-1+ky lives in input parity class
//                         py = (ky+1)&1 at coarse row oy + {-1,0,0,+1}[ky]
//   convT k4 s2 p1 fprop: 4 problems (output parity classes), 4 taps each
//                         qy=0: ky=1 -> iy=m, ky=3 -> iy=m-1 ; qy=1: ky=0 -> iy=m+1, ky=2 -> iy=m
//   conv k4 s2 dgrad == convT fprop form ; convT dgrad == conv k4 s2 fprop form.
#pragma once
#include <stdint.h>

namespace svrs {

struct Tap {
    long long in_off;  // element offset of this tap's input view (parity class base)
    long long w_off;   // element offset of W_t inside the packed weight buffer (or tap id for wgrad)
    int dy, dx;
};

struct Prob {
    long long out_off;  // element offset of this problem's output view
    int ntaps;
    int pad_;
    Tap taps[16];
};

struct TapGeom {
    long long o_sn, o_sy, o_sx;  // output view strides (elements)
    long long i_sn, i_sy, i_sx;  // input view strides (elements)
    int N, OH, OW, IH, IW;
    int K;   // reduction channels (channels of the input view)
    int Nc;  // output channels
    int nprob;
    Prob prob[4];
};

// Builders (host).  Cin/Cout below are the channels of the tensor being READ / WRITTEN by this GEMM.
static inline void geom_conv3(TapGeom& g, int N, int H, int W, int Cr, int Cw, bool flip) {
    g.N = N; g.OH = H; g.OW = W; g.IH = H; g.IW = W; g.K = Cr; g.Nc = Cw; g.nprob = 1;
    g.o_sn = (long long)H * W * Cw; g.o_sy = (long long)W * Cw; g.o_sx = Cw;
    g.i_sn = (long long)H * W * Cr; g.i_sy = (long long)W * Cr; g.i_sx = Cr;
    Prob& p = g.prob[0];
    p.out_off = 0; p.ntaps = 9;
    for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
            Tap& t = p.taps[ky * 3 + kx];
            t.in_off = 0;
            t.w_off = (long long)(ky * 3 + kx) * Cr * Cw;
            t.dy = flip ? 1 - ky : ky - 1;
            t.dx = flip ? 1 - kx : kx - 1;
        }
}

// strided conv form: reads a fine tensor [N,H,W,Cr] through its 4 parity classes, writes coarse [N,H/2,W/2,Cw]
static inline void geom_conv4s2(TapGeom& g, int N, int H, int W, int Cr, int Cw) {
    g.N = N; g.OH = H / 2; g.OW = W / 2; g.IH = H / 2; g.IW = W / 2; g.K = Cr; g.Nc = Cw; g.nprob = 1;
    g.o_sn = (long long)(H / 2) * (W / 2) * Cw; g.o_sy = (long long)(W / 2) * Cw; g.o_sx = Cw;
    g.i_sn = (long long)H * W * Cr; g.i_sy = 2LL * W * Cr; g.i_sx = 2LL * Cr;
    Prob& p = g.prob[0];
    p.out_off = 0; p.ntaps = 16;
    const int par[4] = {1, 0, 1, 0};
    const int sh[4] = {-1, 0, 0, 1};
    for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx) {
            Tap& t = p.taps[ky * 4 + kx];
            t.in_off = ((long long)par[ky] * W + par[kx]) * Cr;
            t.w_off = (long long)(ky * 4 + kx) * Cr * Cw;
            t.dy = sh[ky];
            t.dx = sh[kx];
        }
}

// transposed form: reads coarse [N,H,W,Cr], writes fine [N,2H,2W,Cw] one output parity class per problem
static inline void geom_convT4s2(TapGeom& g, int N, int H, int W, int Cr, int Cw) {
    g.N = N; g.OH = H; g.OW = W; g.IH = H; g.IW = W; g.K = Cr; g.Nc = Cw; g.nprob = 4;
    g.o_sn = 4LL * H * W * Cw; g.o_sy = 4LL * W * Cw; g.o_sx = 2LL * Cw;
    g.i_sn = (long long)H * W * Cr; g.i_sy = (long long)W * Cr; g.i_sx = Cr;
    // for output parity q: the two contributing kernel rows and their input shifts
    const int kk[2][2] = {{1, 3}, {0, 2}};
    const int sh[2][2] = {{0, -1}, {1, 0}};
    for (int qy = 0; qy < 2; ++qy)
        for (int qx = 0; qx < 2; ++qx) {
            Prob& p = g.prob[qy * 2 + qx];
            p.out_off = ((long long)qy * 2 * W + qx) * Cw;
            p.ntaps = 4;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    Tap& t = p.taps[a * 2 + b];
                    int ky = kk[qy][a], kx = kk[qx][b];
                    t.in_off = 0;
                    t.w_off = (long long)(ky * 4 + kx) * Cr * Cw;
                    t.dy = sh[qy][a];
                    t.dx = sh[qx][b];
                }
        }
}

}  // namespace svrs
