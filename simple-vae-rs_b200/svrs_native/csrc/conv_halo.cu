// conv_halo.cu - halo-reuse variant of the tcgen05 implicit-GEMM kernel for the 3x3 stride-1 forms (conv3 fprop and
// dgrad) on maps that tile into 8 x 16 pixel blocks.
//
// conv_tc.cu fetches one shifted 128-pixel A box PER TAP (9 x 16 KB per 64-channel chunk and tile), which makes the
// 64/128-channel layers at 32x32 / 64x64 L2-bandwidth bound.  Here the activation HALO of a tile
// (18 rows x 16 pixels x 64 channels, one TMA box, 36 KB) is loaded ONCE per channel chunk and the nine taps are nine
// shared-memory matrix descriptors into that same tile:
//     tile = 8 (W) x 16 (H) output pixels; M row r = (iy = r / 8, ix = r % 8)
//     tap (dy, dx): descriptor start = halo + ((dy+1) * 16 + (dx+1)) * 128 B, 8-row groups 2048 B apart (one halo row)
// The start address is 128-B but not 1024-B aligned.  MEASURED on B200: the UMMA applies the 128B-swizzle XOR to the
// ABSOLUTE shared-memory address bits (same function TMA used when writing), so such a start needs NO descriptor
// base-offset; filling the base-offset field with (start >> 7) & 7 produced wrong results (first revision of this kernel).
// Weights are either RESIDENT in shared memory for the whole kernel (9 * Cin/64 * n_tile * 128 B <= 96 KB, e.g. the
// 64->64 layers) or streamed through a 4-slot ring of (chunk, tap) slices.
// Warp roles / TMEM double buffering / epilogue are those of conv_tc.cu.
#include "common.cuh"
#include <stdlib.h>
#include "taps.cuh"
#include <cuda.h>
#include <string.h>
#include "tc_ptx.cuh"

namespace svrs {

constexpr int HL_HALO_BYTES = 18 * 16 * 128;          // 36864
constexpr int HL_HALO_SLOTS = 3;
constexpr int HL_W_SLOT_BYTES = 128 * 128;            // one (chunk, tap) slice for n_tile <= 128
constexpr int HL_W_SLOTS = 4;
constexpr int HL_W_RESIDENT_MAX = 96 * 1024;
constexpr int HL_THREADS = 64 + 32 * TC_EPI_WARPS;
// smem: halo ring | weights (resident region or ring) | barriers
constexpr int HL_SMEM_BYTES = HL_HALO_SLOTS * HL_HALO_BYTES + HL_W_RESIDENT_MAX + 1024 + 256;

struct alignas(64) HaloParams {
    CUtensorMap in_map;     // (C, W, H, N) box (64, 16, 18, 1)
    CUtensorMap w_map;      // (K, Nc, taps) box (64, n_tile, 1)
    __nv_bfloat16* out;
    const float* bias;
    int N, OH, OW, tiles_x, tiles_y;
    int Nc, n_tile, n_tiles, kchunks;
    int act, resident;
    int cw;                         // channels per chunk: 64 / 32 / 16 (SWIZZLE_128B / 64B / 32B rows of 2*cw bytes)
    int hy[9], hx[9], wtap[9];      // tap -> halo offset (dy+1, dx+1) and packed-weight tap index
    float* out2;                    // optional fp32 NCHW-flat second output (EpiRow), per-sample stride out2_ld
    long long out2_ld;
    double* bn_sums;                // optional BatchNorm statistics scratch
};

__device__ __forceinline__ uint64_t halo_desc(uint32_t saddr, int mode) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(2048 >> 4) << 32;                  // SBO: next 8-pixel group = next halo row
    d |= (uint64_t)1 << 46;
    if (mode) d |= (uint64_t)((saddr >> 7) & 7) << 49; // matrix base offset: phase of the 128B-swizzle pattern
    d |= (uint64_t)2 << 61;
    return d;
}

// KSUB = cw / 16 = MMAs (K = 16) per tap and chunk
template <int KSUB, bool EXTRA>
__global__ void __launch_bounds__(HL_THREADS, 1) conv3_halo_kernel(const __grid_constant__ HaloParams p) {
    pdl_entry();
    constexpr uint32_t ROWB = 32u * KSUB;                 // bytes of one pixel's channel chunk
    constexpr uint32_t HALO_BYTES = 18u * 16u * ROWB;     // 36864 / 18432 / 9216 (all multiples of 1024)
    constexpr uint32_t LTYPE = KSUB == 4 ? 2u : (KSUB == 2 ? 4u : 6u);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = smem_base + HL_HALO_SLOTS * HL_HALO_BYTES;
    const uint32_t bar_base = w_base + HL_W_RESIDENT_MAX;
    auto hfull = [&](int s) { return bar_base + 8u * s; };
    auto hempty = [&](int s) { return bar_base + 8u * (HL_HALO_SLOTS + s); };
    auto wfull = [&](int s) { return bar_base + 8u * (2 * HL_HALO_SLOTS + s); };
    auto wempty = [&](int s) { return bar_base + 8u * (2 * HL_HALO_SLOTS + HL_W_SLOTS + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (2 * HL_HALO_SLOTS + 2 * HL_W_SLOTS + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (2 * HL_HALO_SLOTS + 2 * HL_W_SLOTS + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * HL_HALO_SLOTS + 2 * HL_W_SLOTS + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    __shared__ float s_bn[EXTRA ? 2 * EPI_BN_MAXC : 1];
    if (EXTRA && p.bn_sums) epi_bn_zero(s_bn);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.in_map);
        prefetch_tmap(&p.w_map);
        for (int s = 0; s < HL_HALO_SLOTS; ++s) { mbar_init(hfull(s), 1); mbar_init(hempty(s), 1); }
        for (int s = 0; s < HL_W_SLOTS; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();     // prologue above overlaps the previous kernel; nothing before this line touches global memory

    const int tiles_pix = p.tiles_x * p.tiles_y * p.N;
    const int total_tiles = tiles_pix * p.n_tiles;
    const uint32_t w_slice = (uint32_t)p.n_tile * ROWB;      // bytes of one (chunk, tap) weight slice

    // tile order: n-tile outermost so that resident weights are loaded once per n-tile change
    if (warp == 0) {
        if (lane == 0) {
            uint32_t hs = 0, hph = 0, ws = 0, wph = 0, rel_ph = 0;
            int cur_nt = -1;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile / tiles_pix;
                int pt = tile % tiles_pix;
                const int tx = pt % p.tiles_x; pt /= p.tiles_x;
                const int ty = pt % p.tiles_y;
                const int n = pt / p.tiles_y;
                if (p.resident && nt != cur_nt) {
                    // (re)load the whole weight set of this n-tile; a RE-load waits until the MMA warp's commit on
                    // wempty(0) says every MMA that read the previous set has retired
                    if (cur_nt >= 0) { mbar_wait(wempty(0), rel_ph); rel_ph ^= 1u; }
                    mbar_expect_tx(wfull(0), (uint32_t)(9 * p.kchunks) * w_slice);
                    for (int kc = 0; kc < p.kchunks; ++kc)
                        for (int t = 0; t < 9; ++t)
                            tma_load_3d(w_base + (uint32_t)(kc * 9 + t) * w_slice, &p.w_map, wfull(0), kc * p.cw, nt * p.n_tile, p.wtap[t]);
                    cur_nt = nt;
                }
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(hempty(hs), hph ^ 1u);
                    mbar_expect_tx(hfull(hs), HALO_BYTES);
                    tma_load_4d(smem_base + hs * HL_HALO_BYTES, &p.in_map, hfull(hs), kc * p.cw, tx * 8 - 1, ty * 16 - 1, n);
                    if (++hs == HL_HALO_SLOTS) { hs = 0; hph ^= 1u; }
                    if (!p.resident) {
                        for (int t = 0; t < 9; ++t) {
                            mbar_wait(wempty(ws), wph ^ 1u);
                            mbar_expect_tx(wfull(ws), w_slice);
                            tma_load_3d(w_base + ws * HL_W_SLOT_BYTES, &p.w_map, wfull(ws), kc * p.cw, nt * p.n_tile, p.wtap[t]);
                            if (++ws == HL_W_SLOTS) { ws = 0; wph ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {       // a single thread issues every MMA; the loop is kept free of avoidable ALU work (issue-bound)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t a_hi = desc_hi(16u * ROWB, LTYPE);      // next 8-pixel group = next halo line (16 pixels)
            const uint32_t b_hi = desc_hi(8u * ROWB, LTYPE);
            uint32_t aoff[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) aoff[t] = (uint32_t)(p.hy[t] * 16 + p.hx[t]) * (ROWB >> 4);   // (halo pixel * ROWB) >> 4
            const uint32_t w_slice16 = w_slice >> 4;
            uint32_t hs = 0, hph = 0, ws = 0, wph = 0, wres_ph = 0, acc = 0, acc_phase = 0;
            int cur_nt = -1;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile / tiles_pix;
                if (p.resident && nt != cur_nt) {
                    if (cur_nt >= 0) tc_commit(wempty(0));   // previous weight set is free once issued MMAs retire
                    mbar_wait(wfull(0), wres_ph);
                    tc_fence_after();
                    wres_ph ^= 1u;
                    cur_nt = nt;
                }
                mbar_wait(tempty(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256u;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(hfull(hs), hph);
                    tc_fence_after();
                    const uint32_t a0 = desc_lo(smem_base + hs * HL_HALO_BYTES, 16u);
                    if (p.resident) {
                        const uint32_t b0 = desc_lo(w_base, 16u) + (uint32_t)(kc * 9) * w_slice16;
#pragma unroll
                        for (int t = 0; t < 9; ++t)
#pragma unroll
                            for (int k = 0; k < KSUB; ++k)
                                tc_mma_lohi(d_tmem, a0 + aoff[t] + 2u * k, a_hi, b0 + (uint32_t)t * w_slice16 + 2u * k, b_hi, idesc,
                                            (t | k) ? 1u : (uint32_t)(kc != 0));
                    } else {
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            mbar_wait(wfull(ws), wph);
                            tc_fence_after();
                            const uint32_t b0 = desc_lo(w_base + ws * HL_W_SLOT_BYTES, 16u);
#pragma unroll
                            for (int k = 0; k < KSUB; ++k)
                                tc_mma_lohi(d_tmem, a0 + aoff[t] + 2u * k, a_hi, b0 + 2u * k, b_hi, idesc,
                                            (t | k) ? 1u : (uint32_t)(kc != 0));
                            tc_commit(wempty(ws));
                            if (++ws == HL_W_SLOTS) { ws = 0; wph ^= 1u; }
                        }
                    }
                    tc_commit(hempty(hs));
                    if (kc == p.kchunks - 1) tc_commit(tfull(acc));
                    if (++hs == HL_HALO_SLOTS) { hs = 0; hph ^= 1u; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        const int q = warp % 4;
        const int half = (warp - 2) / 4;
        const int r = q * 32 + lane;
        const int ix = r % 8, iy = r / 8;
        uint32_t acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile / tiles_pix;
            int pt = tile % tiles_pix;
            const int tx = pt % p.tiles_x; pt /= p.tiles_x;
            const int ty = pt % p.tiles_y;
            const int n = pt / p.tiles_y;
            const int c_base = nt * p.n_tile;
            __nv_bfloat16* orow = p.out + (((long long)n * p.OH + ty * 16 + iy) * p.OW + tx * 8 + ix) * p.Nc + c_base;
            mbar_wait(tfull(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * 256u + ((uint32_t)(q * 32) << 16);
            EpiRow er;
            er.o2 = nullptr; er.hw = 0; er.sbn = nullptr;
            if (EXTRA) {
                er.o2 = p.out2 ? p.out2 + (long long)n * p.out2_ld + (long long)(ty * 16 + iy) * p.OW + (tx * 8 + ix) : nullptr;
                er.hw = p.OH * p.OW;
                er.sbn = p.bn_sums ? s_bn : nullptr;
            }
            epi_dispatch<EXTRA>(p.act, taddr, p.n_tile, half, (!EXTRA || p.out) ? orow : nullptr, p.bias, c_base, p.Nc, true, er);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (EXTRA && p.bn_sums) epi_bn_flush(s_bn, p.bn_sums, p.Nc, threadIdx.x - 64, 32 * TC_EPI_WARPS);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Halo kernel for the TRANSPOSED stride-2 form (ConvTranspose k4 s2 p1 fprop, and dgrad of the k4 s2 conv): the coarse
// input tile (8 x 16 pixels + 1-pixel halo, same 18 x 16 box as above) is loaded once per channel chunk and serves all
// FOUR output-parity sub-convolutions (2 x 2 taps each, taps.cuh geom_convT4s2) = 16 (parity, tap) descriptor views;
// conv_tc.cu re-fetches a shifted A box for every one of them.  Four accumulators (one per output parity, n_tile <= 64
// columns each) live in one 256-column TMEM buffer, double-buffered; the epilogue scatters each parity to its
// stride-2 view of the fine output.  Weights ((chunk, parity, tap) slices of n_tile x cw) stream through a 4-slot ring.
// ------------------------------------------------------------------------------------------------------------------
struct alignas(64) HaloTParams {
    CUtensorMap in_map;     // (C, W, H, N) box (cw, 16, 18, 1) over the coarse input
    CUtensorMap w_map;      // (K, Nc, 16 taps) box (cw, n_tile, 1)
    __nv_bfloat16* out;
    const float* bias;
    int N, H, W, tiles_x, tiles_y;      // coarse dims
    int Nc, n_tile, n_tiles, kchunks, act, cw;
    long long o_sn, o_sy, o_sx;         // element strides of one output-parity view of the fine tensor
    long long out_off[4];               // element offset of each parity view
    int hy[16], hx[16], wtap[16];       // index = parity * 4 + tap
    double* bn_sums;                    // optional BatchNorm statistics scratch (SVRS_BN_REPLICAS x double[2*Nc])
};

template <int KSUB, bool EXTRA>
__global__ void __launch_bounds__(HL_THREADS, 1) convT_halo_kernel(const __grid_constant__ HaloTParams p) {
    pdl_trigger();
    constexpr uint32_t ROWB = 32u * KSUB;
    constexpr uint32_t HALO_BYTES = 18u * 16u * ROWB;
    constexpr uint32_t LTYPE = KSUB == 4 ? 2u : (KSUB == 2 ? 4u : 6u);
    constexpr uint32_t WT_SLOT = 64u * 128u;              // n_tile <= 64 rows of <= 128 B
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = smem_base + HL_HALO_SLOTS * HL_HALO_BYTES;
    const uint32_t bar_base = w_base + HL_W_RESIDENT_MAX;
    auto hfull = [&](int s) { return bar_base + 8u * s; };
    auto hempty = [&](int s) { return bar_base + 8u * (HL_HALO_SLOTS + s); };
    auto wfull = [&](int s) { return bar_base + 8u * (2 * HL_HALO_SLOTS + s); };
    auto wempty = [&](int s) { return bar_base + 8u * (2 * HL_HALO_SLOTS + HL_W_SLOTS + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (2 * HL_HALO_SLOTS + 2 * HL_W_SLOTS + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (2 * HL_HALO_SLOTS + 2 * HL_W_SLOTS + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * HL_HALO_SLOTS + 2 * HL_W_SLOTS + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    __shared__ float s_bn[EXTRA ? 2 * EPI_BN_MAXC : 1];
    if (EXTRA && p.bn_sums) epi_bn_zero(s_bn);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.in_map);
        prefetch_tmap(&p.w_map);
        for (int s = 0; s < HL_HALO_SLOTS; ++s) { mbar_init(hfull(s), 1); mbar_init(hempty(s), 1); }
        for (int s = 0; s < HL_W_SLOTS; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();

    const int tiles_pix = p.tiles_x * p.tiles_y * p.N;
    const int total_tiles = tiles_pix * p.n_tiles;
    const uint32_t w_slice = (uint32_t)p.n_tile * ROWB;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t hs = 0, hph = 0, ws = 0, wph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile / tiles_pix;
                int pt = tile % tiles_pix;
                const int tx = pt % p.tiles_x; pt /= p.tiles_x;
                const int ty = pt % p.tiles_y;
                const int n = pt / p.tiles_y;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(hempty(hs), hph ^ 1u);
                    mbar_expect_tx(hfull(hs), HALO_BYTES);
                    tma_load_4d(smem_base + hs * HL_HALO_BYTES, &p.in_map, hfull(hs), kc * p.cw, tx * 8 - 1, ty * 16 - 1, n);
                    if (++hs == HL_HALO_SLOTS) { hs = 0; hph ^= 1u; }
                    for (int i = 0; i < 16; ++i) {
                        mbar_wait(wempty(ws), wph ^ 1u);
                        mbar_expect_tx(wfull(ws), w_slice);
                        tma_load_3d(w_base + ws * WT_SLOT, &p.w_map, wfull(ws), kc * p.cw, nt * p.n_tile, p.wtap[i]);
                        if (++ws == HL_W_SLOTS) { ws = 0; wph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t a_hi = desc_hi(16u * ROWB, LTYPE);
            const uint32_t b_hi = desc_hi(8u * ROWB, LTYPE);
            uint32_t aoff[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) aoff[i] = (uint32_t)(p.hy[i] * 16 + p.hx[i]) * (ROWB >> 4);
            uint32_t hs = 0, hph = 0, ws = 0, wph = 0, acc = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tempty(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256u;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(hfull(hs), hph);
                    tc_fence_after();
                    const uint32_t a0 = desc_lo(smem_base + hs * HL_HALO_BYTES, 16u);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {              // i = parity * 4 + tap
                        mbar_wait(wfull(ws), wph);
                        tc_fence_after();
                        const uint32_t b0 = desc_lo(w_base + ws * WT_SLOT, 16u);
#pragma unroll
                        for (int k = 0; k < KSUB; ++k)
                            tc_mma_lohi(d_tmem + (uint32_t)((i >> 2) * p.n_tile), a0 + aoff[i] + 2u * k, a_hi, b0 + 2u * k, b_hi, idesc,
                                        ((i & 3) | k) ? 1u : (uint32_t)(kc != 0));
                        tc_commit(wempty(ws));
                        if (++ws == HL_W_SLOTS) { ws = 0; wph ^= 1u; }
                    }
                    tc_commit(hempty(hs));
                    if (kc == p.kchunks - 1) tc_commit(tfull(acc));
                    if (++hs == HL_HALO_SLOTS) { hs = 0; hph ^= 1u; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        const int q = warp % 4;
        const int half = (warp - 2) / 4;
        const int r = q * 32 + lane;
        const int ix = r % 8, iy = r / 8;
        uint32_t acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile / tiles_pix;
            int pt = tile % tiles_pix;
            const int tx = pt % p.tiles_x; pt /= p.tiles_x;
            const int ty = pt % p.tiles_y;
            const int n = pt / p.tiles_y;
            const int c_base = nt * p.n_tile;
            const long long pix = (long long)n * p.o_sn + (long long)(ty * 16 + iy) * p.o_sy + (long long)(tx * 8 + ix) * p.o_sx + c_base;
            mbar_wait(tfull(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            EpiRow er;
            er.o2 = nullptr; er.hw = 0;
            er.sbn = (EXTRA && p.bn_sums) ? s_bn : nullptr;
            for (int par = 0; par < 4; ++par) {
                const uint32_t taddr = tmem_base + acc * 256u + (uint32_t)(par * p.n_tile) + ((uint32_t)(q * 32) << 16);
                epi_dispatch<EXTRA>(p.act, taddr, p.n_tile, half, p.out + p.out_off[par] + pix, p.bias, c_base, p.Nc, true, er);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (EXTRA && p.bn_sums) epi_bn_flush(s_bn, p.bn_sums, p.Nc, threadIdx.x - 64, 32 * TC_EPI_WARPS);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// persistent-grid cap of the halo kernels (SVRS_HALO_MAX_CTAS, default = all SMs)
static int halo_max_ctas() {
    static const int v = getenv("SVRS_HALO_MAX_CTAS") ? atoi(getenv("SVRS_HALO_MAX_CTAS")) : num_sms();
    return v < 1 ? 1 : v;
}
static int g_halo_mode = 1;   // 0 = off (per-tap TMA kernels), 1 = on
void set_halo_mode(int m) { g_halo_mode = m != 0; }
int get_halo_mode() { return g_halo_mode; }
int make_w_map_pub(CUtensorMap* m, const void* base, int K, int Nc, int taps, int n_tile, int cw);

bool halo_supported(int form, int Cr, int Cw, int OW, int OH) {
    const bool cr_ok = Cr % 64 == 0 || Cr == 32 || Cr == 16;      // one 16/32-channel chunk, or 64-channel chunks
    return g_halo_mode != 0 && (form == 0 || form == 1) && cr_ok && Cw % 16 == 0 && Cw >= 16 && OW % 8 == 0 && OH % 16 == 0;
}

bool convT_halo_supported(int Cr, int Cw, int W, int H) {      // W, H: coarse (input) dims
    const bool cr_ok = Cr % 64 == 0 || Cr == 32 || Cr == 16;
    const bool cw_ok = Cw % 16 == 0 && Cw >= 16 && (Cw <= 64 || Cw % 64 == 0);
    return g_halo_mode != 0 && cr_ok && cw_ok && W % 8 == 0 && H % 16 == 0;
}

int launch_convT_halo(const void* in, const void* w_nk, const float* bias, void* out, int N, int H, int W, int Cr, int Cw, int act,
                      const ConvExtra& ex, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaSuccess;
#define SET_T(K, X) if (e == cudaSuccess) e = cudaFuncSetAttribute(convT_halo_kernel<K, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM_BYTES)
        SET_T(4, false); SET_T(2, false); SET_T(1, false); SET_T(4, true); SET_T(2, true); SET_T(1, true);
#undef SET_T
        if (e != cudaSuccess) { set_error("convT_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SVRS_E_CUDA; }
        attr_set = true;
    }
    TapGeom g;
    geom_convT4s2(g, N, H, W, Cr, Cw);
    HaloTParams p;
    memset(&p, 0, sizeof(p));
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.bias = bias;
    p.N = N; p.H = H; p.W = W; p.tiles_x = W / 8; p.tiles_y = H / 16;
    p.Nc = Cw;
    p.n_tile = Cw <= 64 ? Cw : 64;
    p.n_tiles = Cw / p.n_tile;
    p.cw = Cr % 64 == 0 ? 64 : Cr;
    p.kchunks = Cr / p.cw;
    p.act = act;
    p.bn_sums = ex.bn_sums;
    p.o_sn = g.o_sn; p.o_sy = g.o_sy; p.o_sx = g.o_sx;
    for (int q = 0; q < 4; ++q) {
        p.out_off[q] = g.prob[q].out_off;
        for (int t = 0; t < 4; ++t) {
            const Tap& tp = g.prob[q].taps[t];
            p.hy[q * 4 + t] = tp.dy + 1;
            p.hx[q * 4 + t] = tp.dx + 1;
            p.wtap[q * 4 + t] = (int)(tp.w_off / ((long long)Cr * Cw));
        }
    }
    int rc = make_act_map(&p.in_map, in, Cr, W, H, N, Cr, (long long)W * Cr, (long long)H * W * Cr, 16, 18, 1, p.cw);
    if (rc) return rc;
    rc = make_w_map_pub(&p.w_map, w_nk, Cr, Cw, 16, p.n_tile, p.cw);
    if (rc) return rc;
    long long total = (long long)p.tiles_x * p.tiles_y * N * p.n_tiles;
    int grid = (int)(total < halo_max_ctas() ? total : halo_max_ctas());
    if (grid < 1) return 0;
    const bool extra = ex.bn_sums != nullptr;
#define GO_T(K) do { if (extra) SVRS_LAUNCH((convT_halo_kernel<K, true>), grid, HL_THREADS, HL_SMEM_BYTES, st, p); \
                     else SVRS_LAUNCH((convT_halo_kernel<K, false>), grid, HL_THREADS, HL_SMEM_BYTES, st, p); } while (0)
    if (p.cw == 64) GO_T(4); else if (p.cw == 32) GO_T(2); else GO_T(1);
#undef GO_T
    return check_launch("convT_halo_kernel");
}

int launch_conv3_halo(int form, const void* in, const void* w_nk, const float* bias, void* out, int N, int H, int W, int Cr, int Cw,
                      int act, const ConvExtra& ex, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaSuccess;
#define SET_C(K, X) if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3_halo_kernel<K, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM_BYTES)
        SET_C(4, false); SET_C(2, false); SET_C(1, false); SET_C(4, true); SET_C(2, true); SET_C(1, true);
#undef SET_C
        if (e != cudaSuccess) { set_error("conv3_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SVRS_E_CUDA; }
        attr_set = true;
    }
    HaloParams p;
    memset(&p, 0, sizeof(p));
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.bias = bias;
    p.N = N; p.OH = H; p.OW = W; p.tiles_x = W / 8; p.tiles_y = H / 16;
    p.Nc = Cw;
    p.n_tile = Cw <= 128 ? Cw : 128;
    p.n_tiles = (Cw + p.n_tile - 1) / p.n_tile;
    p.cw = Cr % 64 == 0 ? 64 : Cr;
    p.kchunks = Cr / p.cw;
    p.act = act;
    p.out2 = ex.out2; p.out2_ld = ex.out2_ld; p.bn_sums = ex.bn_sums;
    p.resident = (9 * p.kchunks * p.n_tile * 2 * p.cw <= HL_W_RESIDENT_MAX) ? 1 : 0;
    for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
            int t = ky * 3 + kx;
            int dy = form == 1 ? 1 - ky : ky - 1, dx = form == 1 ? 1 - kx : kx - 1;
            p.hy[t] = dy + 1; p.hx[t] = dx + 1; p.wtap[t] = t;
        }
    // halo box: 16 pixels wide (x0-1 .. x0+14), 18 rows (y0-1 .. y0+16), one image
    int rc = make_act_map(&p.in_map, in, Cr, W, H, N, Cr, (long long)W * Cr, (long long)H * W * Cr, 16, 18, 1, p.cw);
    if (rc) return rc;
    rc = make_w_map_pub(&p.w_map, w_nk, Cr, Cw, 9, p.n_tile, p.cw);
    if (rc) return rc;
    long long total = (long long)p.tiles_x * p.tiles_y * N * p.n_tiles;
    int grid = (int)(total < halo_max_ctas() ? total : halo_max_ctas());
    if (grid < 1) return 0;
    const bool extra = ex.bn_sums != nullptr || ex.out2 != nullptr;
#define GO_C(K) do { if (extra) SVRS_LAUNCH((conv3_halo_kernel<K, true>), grid, HL_THREADS, HL_SMEM_BYTES, st, p); \
                     else SVRS_LAUNCH((conv3_halo_kernel<K, false>), grid, HL_THREADS, HL_SMEM_BYTES, st, p); } while (0)
    if (p.cw == 64) GO_C(4); else if (p.cw == 32) GO_C(2); else GO_C(1);
#undef GO_C
    return check_launch("conv3_halo_kernel");
}

}  // namespace svrs
