// Shared pieces of the tcgen05 kernels (conv_tc.cu, conv_tma.cu): tile shape, K/N chunk plan, PTX wrappers, UMMA
// descriptors, tile iterator.
#pragma once
#include "common.cuh"
#include <stdlib.h>

namespace semb {

constexpr int TILE_H = 16, TILE_W = 8;          // 128 output pixels = UMMA M
constexpr int TC_THREADS = 128;

struct TcPlan { int KC, NC, nchunks, kchunks, tmem_cols; };

// Shared-memory budget: A planes + B taps must leave room for 2 CTAs per SM.
static inline TcPlan tc_plan_nc(int Cin, int Cout, int taps, int nc_max, int a_bytes_per_channel) {
    TcPlan p;
    const int c16 = (Cout + 15) / 16 * 16;
    p.nchunks = (c16 + nc_max - 1) / nc_max;
    p.NC = ((c16 + p.nchunks - 1) / p.nchunks + 15) / 16 * 16;
    const int cin16 = (Cin + 15) / 16 * 16;
    int kc = 64;
    static const size_t stage_budget = [] { const char* e = getenv("SEMB_TC_STAGE_KB"); const int v = e ? atoi(e) : 96; return (size_t)(v >= 16 && v <= 200 ? v : 96) * 1024; }();
    while (kc > 16 && (size_t)taps * kc * p.NC * 2 + (size_t)kc * a_bytes_per_channel > stage_budget) kc >>= 1;
    if (kc > cin16) kc = cin16 <= 16 ? 16 : (cin16 <= 32 ? 32 : 64);
    p.KC = kc;
    p.kchunks = (Cin + kc - 1) / kc;
    p.tmem_cols = p.NC <= 32 ? 32 : (p.NC <= 64 ? 64 : (p.NC <= 128 ? 128 : 256));
    return p;
}

// The plan is a function of (Cin, Cout, taps) only: weight packing and every launch agree on it without sharing state.
// Layers whose weights do not fit one K chunk (streamed per chunk from L2) are planned with N <= 128, which is what the
// two-tiles-per-weight-chunk mode of conv_tma.cu needs (two tiles x two accumulator generations in 512 TMEM columns;
// M128 x N128 x K16 is also where the A-operand fetch and the MMA take the same 64 cycles).
static inline TcPlan tc_plan(int Cin, int Cout, int taps) {
    static const int nc_max = [] { const char* e = getenv("SEMB_TC_NC_MAX"); const int v = e ? atoi(e) : 256; return v >= 16 && v <= 256 ? v / 16 * 16 : 256; }();
    static const bool pair_on = [] { const char* e = getenv("SEMB_TMA_NO_PAIR"); return !(e && e[0] && e[0] != '0'); }();
    TcPlan p = tc_plan_nc(Cin, Cout, taps, nc_max, 362);            // 16x8 tile: 18 x 10 halo pixels x 2 B, + slack
    // streamed: N <= 128 and the K chunk sized for the 18 x 18 halo of a tile pair (5248 B per 8-channel plane)
    if (pair_on && p.kchunks > 1) p = tc_plan_nc(Cin, Cout, taps, nc_max < 128 ? nc_max : 128, 656);
    return p;
}

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // Plain try_wait spin (measured: a suspend-time hint or a __nanosleep back-off make the hand-offs slower and the
    // kernels are hand-off-latency bound).  Bounded: a descriptor bug must trap, not hang the GPU box.
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();
    }
}
// warp-collective wait: one lane polls, the warp re-converges on it
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// descriptor passed as (lo, hi) halves: only the low word (start address) changes between MMAs of a kernel
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
// Warp-convergent issue (all 32 lanes execute the call, elect.sync picks the issuing lane).  Issuing from inside an
// `if (lane == 0)` region makes ptxas wrap every UTCHMMA / UTCBAR / UTMALDG in a divergence "waterfall" loop
// (ELECT + BRA.U.ANY + R2UR, ~100 cycles per MMA: measured in round 1 as 3500 cycles per 3x3 tile with all data
// movement switched off); with the warp converged the operands stay in uniform registers.
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, majors, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;      // src-size 0 -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// asynchronous arrive: fires when all cp.async issued so far by this thread have landed (no wait in the producer)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}


__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Sum 16 values across the 32 lanes of a warp.  Each round exchanges half of the remaining values with the partner
// lane (offset 16, 8, 4, 2) and keeps the other half; a last round (offset 1) adds the two partial totals.  Result:
// w[0] of lane l holds the warp total of value index (l >> 1) & 15 (same value in lanes l and l^1).  Fixed order.
__device__ __forceinline__ void warp_reduce16(float (&w)[16], int lane) {
#pragma unroll
    for (int h = 8, o = 16; h >= 1; h >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? w[i] : w[i + h];
            const float keep = up ? w[i + h] : w[i];
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
}

struct TcArgs {
    int N, H, W, OH, OW, Cin, Cout, R, S, pad_t, pad_l, pad_mode;
    const bf16* x; int x_pitch, x_coff;
    bf16* y; int y_pitch, y_coff;
    const bf16* wp; const float* bias;
    double* stats; int stats_nstride, stats_cstride;
    int accumulate;
    int tiles_x, tiles_y, total_tiles;
    TcPlan p;
    int plane_bytes, halo_h, halo_w;
    int stages, b_resident, a_bytes, b_bytes, stage_bytes;
    int dbg;      // SEMB_TC_DEBUG ablation bits (profiling only): 1 no A loads, 2 no MMAs, 4 no stores, 8 no moments
    int out_f32;  // y is an fp32 tensor (parity mode: bf16 x 3 split operands, fp32 results); y_pitch / y_coff count floats
    // depth-to-space epilogue (Conv2DTranspose 2x2 stride 2 as a 1x1 conv with 4*C outputs): d2s_c = C > 0 stores output channel
    // (2r+s)*C + c of pixel (y, x) at pixel (2y+r, 2x+s), channel c of a (N, d2s_h, d2s_w) tensor; bias is C long
    int d2s_c, d2s_h, d2s_w;
};

// Walks the tiles t = first + i*stride of an (N, tiles_y, tiles_x) grid without integer divisions in the loop.
struct TileIter {
    int n, ty, tx, dn, dty, dtx, tiles_x, tiles_y;
    __device__ __forceinline__ TileIter(int first, int stride, int tiles_x_, int tiles_y_) : tiles_x(tiles_x_), tiles_y(tiles_y_) {
        const int tpi = tiles_x * tiles_y;
        n = first / tpi;
        int rem = first - n * tpi;
        ty = rem / tiles_x;
        tx = rem - ty * tiles_x;
        dn = stride / tpi;
        rem = stride - dn * tpi;
        dty = rem / tiles_x;
        dtx = rem - dty * tiles_x;
    }
    __device__ __forceinline__ void next() {
        tx += dtx; ty += dty; n += dn;
        if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        if (ty >= tiles_y) { ty -= tiles_y; ++n; }
    }
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// NCT > 0: all output channels fit NCT (16 or 32) accumulator columns -- the HBM-bound high-resolution layers.  Their
// epilogue keeps the per-channel moments of a thread's pixel in REGISTERS across all tiles of the CTA and reduces across
// lanes once per CTA (or per sample), instead of 16 shuffles per 8 channels per tile.  NCT == 0: generic path.

// conv_tma.cu: TMA-staged variant of the forward / data-gradient conv for zero-padded geometries
int conv_tma_launch(const semb_conv_geom* g, const semb_tensor* x, const void* w_packed, const float* bias, const semb_tensor* y,
                    void* stats, int32_t stats_nstride, int32_t stats_cstride, int32_t accumulate, void* stream, int out_f32 = 0,
                    int d2s_h = 0, int d2s_w = 0);

// wgrad_tma.cu: TMA-staged weight gradient of the zero-padded 3x3 layers
int wgrad_tma_launch(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy, float* dw, void* workspace, void* stream);
size_t wgrad_tma_workspace_bytes(const semb_conv_geom* g);

}  // namespace semb
