// Zero-padded stride-1 3x3 / 1x1 convolution (forward and data gradient) on tcgen05 with TMA-staged halo tiles.
//
// Same implicit GEMM as conv_tc.cu (M = 16x8 output pixels, N = Cout, K = taps x Cin, halo tile staged once as
// channel-group planes, taps = shifted UMMA descriptors), but the planes are written by the TMA unit: one elected
// thread issues ONE cp.async.bulk.tensor per 8-channel plane -- a 4-D box {8 channels, halo_w, halo_h, 1} of the NHWC
// tensor, whose dense shared-memory image [halo_h][halo_w][8] IS the no-swizzle K-major core-matrix layout -- and
// the zero padding is the tensor map's out-of-bounds fill.  Round-1 ablation (scripts/ablate_conv.sh) showed the
// cp.async version to be bound by the per-pixel address arithmetic of its 128 producer threads and by its single
// epilogue warpgroup (~77 us for a 256x256x32 batch with loads, MMAs and stores all switched off), not by HBM.
//
//   warps 0-3  epilogue group 0 : even tiles of the CTA
//   warps 4-7  epilogue group 1 : odd tiles
//   warps 8-9  MMA issuers      : even / odd tiles (ncu, round 1: ONE issuing warp needs ~280 instructions = 2100 cycles
//                                 per 3x3 tile and never waits -- it, not HBM or the tensor pipe, paced the kernel)
//   warp  10   TMA producer
//
// Replaces F.conv2d / its data gradient under keras.layers.Conv2D (UNet_Segmentation.py:421,465-468,490-499).
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace semb {

constexpr int TM_THREADS = 352;
constexpr int TM_MAX_STAGES = 6;

constexpr int TM_MAX_ACC = 4;

struct TmArgs {
    TcArgs t;
    int tmem_cols;
    int nmma;             // MMA issuer warps in use: 2 (ring split in halves) when the ring has >= 4 stages, else 1
    int nacc;             // accumulator buffers in TMEM (2 or 4): the MMA -> epilogue -> MMA hand-off takes ~2-4k cycles
    int plane_pitch;      // bytes between channel-group planes of a stage (multiple of 128)
    int plane_box;        // bytes the TMA writes per plane (halo_h * halo_w * 16)
    int nstages;
    int bsplit;           // bulk copies per streamed weight chunk (b_bytes / bsplit must be a multiple of 16)
};

// all three are called by a CONVERGED warp; elect.sync picks the lane that issues (see umma_bf16_elect)
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n\t}"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc_rt(uint32_t slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// NCT: compile-time channel count of the register-moment epilogue (16, 0 = generic).  KR: kernel size.
// KS: K steps per chunk when all input channels fit one chunk (1..4, straight-line MMA issue), 0 = run-time.
// CS ("column split", NCT = 16 with an MMA width of 32): both epilogue groups work on EVERY tile, group g on the accumulator
// columns [16 g, 16 g + 16), so that 32-channel layers (the merged res_path convs, the 1x1 shortcuts of the full-resolution
// blocks) keep their moments in registers too -- the generic epilogue's per-tile butterfly reductions made N = 32 cost
// 92 us where N = 16 cost 49 (256x256, batch 32).
// PAIR (streamed weights only, NCT = KS = 0): the CTA walks 16x16 SUPER-tiles (two 16x8 tiles side by side, one 18x18 halo per
// channel-group plane).  Every stage carries ONE weight chunk that both tiles use -- MMA warp j issues tile j's MMAs from the same
// stage into its own accumulator, epilogue group j stores tile j -- so the L2 -> shared-memory weight stream, which paced the
// deep layers (CycleGAN 512->512 at 32x32: 2.4 MB of weights per 128 output pixels, 88 us against a 19 us tensor-pipe floor), is
// halved per output pixel, and the chunk plan caps N at 128 (A fetch and MMA both 64 cycles) so that two tiles x two
// accumulator generations fit the 512 TMEM columns.
template <int NCT, int KR, int KS, bool CS = false, bool PAIR = false>
__global__ void __launch_bounds__(TM_THREADS, PAIR ? 1 : 2) conv_tma_kernel(const TmArgs args, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TM_MAX_STAGES], empty_bar[TM_MAX_STAGES], acc_full[TM_MAX_ACC], acc_empty[TM_MAX_ACC], b_full;
    __shared__ uint32_t tmem_slot;
    pdl_trigger();                  // the successor's prologue may overlap this kernel's tail
    const TcArgs& a = args.t;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NCM = CS ? 32 : NCT;                 // MMA width when it is a compile-time constant
    constexpr int PN = CS ? 16 : NCT;                  // channels of one epilogue group (register-moment path)
    const int KC = a.p.KC, NC = NCT > 0 ? NCM : a.p.NC, kchunks = a.p.kchunks;
    const int nchunk = blockIdx.y;
    const int nstages = args.nstages;
    const int nacc = args.nacc;
    constexpr int TW = PAIR ? 2 * TILE_W : TILE_W;     // width of the (super-)tile one ring stage serves
    constexpr int HALO_W = TW + KR - 1;
    // smem (128-byte aligned): [resident B (optional)] [nstages x (A planes [+ B chunk])] [moment partials 2 x 8 x NC floats]
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring_u32 = smem_base + (a.b_resident ? a.b_bytes : 0);
    float* part_all = reinterpret_cast<float*>(smem + (a.b_resident ? a.b_bytes : 0) + (size_t)nstages * a.stage_bytes);
    const int ntiles = (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 8) tmem_alloc_rt(smem_u32(&tmem_slot), (uint32_t)args.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < nstages; ++i) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), PAIR ? 2 : 1); }
        for (int i = 0; i < TM_MAX_ACC; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), CS ? 256 : 128); }
        mbar_init(smem_u32(&b_full), 1);
    }
    // Planes beyond Cin (Cin % 16 == 8) are never written by the TMA: they must hold finite values (their weights are 0).
    {
        const int n16 = (nstages * a.stage_bytes) >> 4;
        uint4* ring = reinterpret_cast<uint4*>(smem + (a.b_resident ? a.b_bytes : 0));
        for (int i = tid; i < n16; i += TM_THREADS) ring[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < 16 * NC; i += TM_THREADS) part_all[i] = 0.f;
    fence_proxy_async();            // generic-proxy zero fill -> later async-proxy (TMA / tensor core) accesses
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();                     // everything above touched only shared memory / TMEM; global memory from here on

    if (warp == 10) {
        // ============================== TMA producer (whole warp converged) ==============================
        {
            if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            __syncwarp();
            const int taps = KR * KR;
            if (a.b_resident) {
                mbar_expect_tx(smem_u32(&b_full), (uint32_t)a.b_bytes);
                bulk_load(smem_base, a.wp + (size_t)nchunk * kchunks * taps * KC * NC, (uint32_t)a.b_bytes, smem_u32(&b_full));
            }
            // The ring is split in two halves, one per MMA warp (tile parity): every barrier has exactly one consumer
            // that waits on each of its phases in order (a consumer that skipped phases could not tell them apart).
            TileIter it(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
            // (PAIR: one ring, both MMA warps consume every stage and the empty barriers count two commits)
            const bool split = !PAIR && args.nmma == 2;
            const int half = split ? nstages >> 1 : nstages;
            int slot0 = 0, slot1 = 0;
            uint32_t ph0 = 1, ph1 = 1;
            for (int ti = 0; ti < ntiles; ++ti, it.next()) {
                const int y0 = it.ty * TILE_H - a.pad_t, x0 = it.tx * TW - a.pad_l;
                const int w = split ? (ti & 1) : 0;
                for (int kc = 0; kc < kchunks; ++kc) {
                    const int stage = w ? half + slot1 : slot0;
                    mbar_wait(smem_u32(&empty_bar[stage]), w ? ph1 : ph0);
                    const int c0 = kc * KC;
                    const int real = min(KC, a.Cin - c0) >> 3;             // planes that exist in the tensor
                    const uint32_t sbase = ring_u32 + stage * a.stage_bytes;
                    const uint32_t bar = smem_u32(&full_bar[stage]);
                    mbar_expect_tx(bar, (uint32_t)(real * args.plane_box + (a.b_resident ? 0 : a.b_bytes)));
                    if (!(a.dbg & 1)) {
                        for (int k8 = 0; k8 < real; ++k8)
                            tma_load_4d(sbase + k8 * args.plane_pitch, &xmap, a.x_coff + c0 + k8 * 8, x0, y0, it.n, bar);
                    } else {
                        // ablation: complete the transaction count without moving data
                        asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                                     "@e mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n\t}" ::"r"(bar), "r"((uint32_t)(real * args.plane_box)) : "memory");
                    }
                    if (!a.b_resident) {
                        // the weight chunk of a stage as `bsplit` concurrent bulk copies (one 70 KB copy is served at ~25 B/clk)
                        const uint32_t piece = (uint32_t)a.b_bytes / (uint32_t)args.bsplit;
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wp + ((size_t)nchunk * kchunks + kc) * taps * KC * NC);
                        for (int q = 0; q < args.bsplit; ++q)
                            bulk_load(sbase + a.a_bytes + q * piece, src + (size_t)q * piece, piece, bar);
                    }
                    if (w) { if (++slot1 == half) { slot1 = 0; ph1 ^= 1; } }
                    else   { if (++slot0 == half) { slot0 = 0; ph0 ^= 1; } }
                }
            }
        }
    } else if (warp >= 8) {
        // ============================== MMA issuers (two warps, alternate tiles; converged, elect.sync per instruction) ====
        if (warp - 8 < args.nmma) {
            const int mw = warp - 8;
            const int nmma = args.nmma;
            const uint32_t idesc = instr_desc(128, NC, 0, 0);
            const uint64_t ad0 = smem_desc(ring_u32, args.plane_pitch, HALO_W * 16);
            const uint64_t bd0 = smem_desc(a.b_resident ? smem_base : ring_u32 + a.a_bytes, NC * 16, 128);
            const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
            const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
            const uint32_t pl2 = 2 * ((uint32_t)args.plane_pitch >> 4);
            const uint32_t nc2 = 2 * (uint32_t)NC;
            const uint32_t stage16 = (uint32_t)a.stage_bytes >> 4;
            const uint32_t btap = (uint32_t)(KC / 8) * NC;
            if (a.b_resident) mbar_wait(smem_u32(&b_full), 0);
            const int half = PAIR ? nstages : (nmma == 2 ? nstages >> 1 : nstages);
            int slot = 0, buf = mw;
            uint32_t phase = 0, aphase = 1;
            for (int ti = PAIR ? 0 : mw; ti < ntiles; ti += PAIR ? 1 : nmma) {
                mbar_wait(smem_u32(&acc_empty[buf]), aphase);
                tc_fence_after();
                const uint32_t dcol = tmem + buf * NC;
                for (int kc = 0; kc < kchunks; ++kc) {
                    const int stage = PAIR ? slot : mw * half + slot;
                    mbar_wait(smem_u32(&full_bar[stage]), phase);
                    tc_fence_after();
                    const int ksteps = KS > 0 ? KS : (min(KC, a.Cin - kc * KC) + 15) / 16;
                    const uint32_t soff = (uint32_t)stage * stage16;
                    const uint32_t a_lo = a_lo0 + soff + (PAIR ? (uint32_t)(mw * TILE_W) : 0u);    // tile j of the pair: 8 pixels = 8 x 16 B to the right
                    const uint32_t b_lo = b_lo0 + (a.b_resident ? 0u : soff);
                    if (!(a.dbg & 2)) {
#pragma unroll
                        for (int tap = 0; tap < KR * KR; ++tap) {
                            const uint32_t at = a_lo + (uint32_t)((tap / KR) * HALO_W + (tap % KR));
                            const uint32_t bt = b_lo + (uint32_t)tap * btap;
#pragma unroll
                            for (int ks = 0; ks < (KS > 0 ? KS : 4); ++ks) {      // KC <= 64: at most four K steps per chunk
                                if (KS > 0 || ks < ksteps)
                                    umma_bf16_elect(dcol, at + ks * pl2, a_hi, bt + ks * nc2, b_hi, idesc,
                                                    (tap | ks) != 0 ? 1u : (uint32_t)(kc != 0));
                            }
                        }
                    }
                    umma_commit_elect(smem_u32(&empty_bar[stage]));
                    if (++slot == half) { slot = 0; phase ^= 1; }
                }
                umma_commit_elect(smem_u32(&acc_full[buf]));
                buf += nmma;                                // this warp's accumulator buffers: mw, mw + nmma, ...
                if (buf >= nacc) { buf = mw; aphase ^= 1; }
            }
        }
    } else {
        // ============================== epilogue (two groups, alternate tiles; accumulator buffer = tile % nacc) ==============================
        const int grp = warp >> 2;                        // 0 or 1: even / odd tiles of the CTA
        const int wq = warp & 3;                          // TMEM lane quarter this warp may read
        const int etid = tid & 127;
        float* part = part_all + grp * 8 * NC;
        const int PW = CS ? 16 : NC;                       // channels this group reduces (row pitch of its partials)
        const int c_begin = nchunk * NC + (CS ? grp * 16 : 0);
        int cur_n = -1;
        auto combine = [&](int n) {
            // the group's 128 threads: combine the four warps' partial moments in a fixed order, one fp64 atomic per channel
            asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
            for (int i = etid; i < PW; i += 128) {
                const int c = c_begin + i;
                if (c < a.Cout) {
                    const float t1 = ((part[i] + part[PW + i]) + part[2 * PW + i]) + part[3 * PW + i];
                    const float t2 = ((part[4 * PW + i] + part[5 * PW + i]) + part[6 * PW + i]) + part[7 * PW + i];
                    double* st = a.stats + (size_t)n * a.stats_nstride + c;
                    atomicAdd(st, (double)t1);
                    atomicAdd(st + a.stats_cstride, (double)t2);
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
            for (int i = etid; i < 8 * PW; i += 128) part[i] = 0.f;
            asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
        };
        const int m = wq * 32 + lane;
        const int my = m / TILE_W, mx = m % TILE_W;
        const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
        const int nuse = nacc >> 1;                        // buffers of this group: grp, grp + 2, ...
        // nacc is 2 or 4 (conv_tma_launch): the per-tile buffer index / phase are masks and shifts -- `ti % nacc` and `k / nuse` by
        // run-time values were a signed-division sequence per tile in all eight epilogue warps (ncu source page: IABS / I2F lines
        // among the top stall sites of the issue-bound full-resolution launches)
        const int amask = nacc - 1, ashift = nacc == 4 ? 2 : 1, umask = nuse - 1, ushift = ashift - 1;
        // (CS and PAIR: both groups visit every (super-)tile of the CTA; otherwise the groups alternate tiles)
        TileIter it((CS || PAIR) ? blockIdx.x : blockIdx.x + grp * gridDim.x, (CS || PAIR) ? gridDim.x : 2 * gridDim.x, a.tiles_x, a.tiles_y);
        if constexpr (NCT > 0) {
            float s1[NCT], s2[NCT];
#pragma unroll
            for (int i = 0; i < NCT; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
            auto flush_regs = [&](int n) {
#pragma unroll
                for (int q = 0; q < NCT / 16; ++q) {
                    float w[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = s1[q * 16 + i];
                    warp_reduce16(w, lane);
                    if ((lane & 1) == 0) part[wq * PN + q * 16 + (lane >> 1)] = w[0];
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = s2[q * 16 + i];
                    warp_reduce16(w, lane);
                    if ((lane & 1) == 0) part[(4 + wq) * PN + q * 16 + (lane >> 1)] = w[0];
                }
#pragma unroll
                for (int i = 0; i < NCT; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
                combine(n);
            };
            const int cg0 = CS ? grp * 16 : 0;                  // first channel (of the chunk) this group stores / reduces
            for (int ti = CS ? 0 : grp, k = 0; ti < ntiles; ti += CS ? 1 : 2, ++k, it.next()) {
                const int buf = ti & amask;
                const uint32_t acc_addr = lane_addr + buf * NC + cg0;
                const uint32_t full_u32 = smem_u32(&acc_full[buf]), empty_u32 = smem_u32(&acc_empty[buf]);
                const uint32_t fpar = CS ? (uint32_t)(ti >> ashift) & 1u : (uint32_t)(k >> ushift) & 1u;
                if (a.stats && a.stats_nstride != 0 && cur_n >= 0 && it.n != cur_n) flush_regs(cur_n);
                cur_n = it.n;
                const int oy = it.ty * TILE_H + my, ox = it.tx * TILE_W + mx;
                const bool pvalid = oy < a.OH && ox < a.OW;
                bf16* yp = a.y + ((size_t)(it.n * a.OH + (pvalid ? oy : 0)) * a.OW + (pvalid ? ox : 0)) * a.y_pitch + a.y_coff;
                float* ypf = reinterpret_cast<float*>(a.y) + ((size_t)(it.n * a.OH + (pvalid ? oy : 0)) * a.OW + (pvalid ? ox : 0)) * a.y_pitch + a.y_coff;
                mbar_wait_warp(full_u32, fpar);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < NCT / 16; ++h) {
                    float v16[16];
                    tmem_ld16(acc_addr + h * 16, v16);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int cl = h * 16 + half * 8;           // channel inside this group's register moments
                        const int c = cg0 + cl;                     // channel of the output tensor
                        const bool cvalid = c < a.Cout;
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = v16[half * 8 + i];
                        if (a.bias && cvalid) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += a.bias[c + i];
                        }
                        if (a.stats && pvalid && !(a.dbg & 8)) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) { s1[cl + i] += v[i]; s2[cl + i] = fmaf(v[i], v[i], s2[cl + i]); }
                        }
                        if (pvalid && cvalid && !(a.dbg & 4)) {
                            if (a.accumulate) {
                                float o[8];
                                if (a.out_f32) Vec8<float>::load(ypf + c, o); else Vec8<bf16>::load(yp + c, o);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] += o[i];
                            }
                            if (a.out_f32) Vec8<float>::store(ypf + c, v); else Vec8<bf16>::store(yp + c, v);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(empty_u32);
            }
            if (a.stats && cur_n >= 0) flush_regs(a.stats_nstride != 0 ? cur_n : 0);
        } else {
            const int d2s_q0 = a.d2s_c ? c_begin / a.d2s_c : 0, d2s_c0 = a.d2s_c ? c_begin - d2s_q0 * a.d2s_c : 0;
            for (int ti = PAIR ? 0 : grp, k = 0; ti < ntiles; ti += PAIR ? 1 : 2, ++k, it.next()) {
                const int buf = PAIR ? grp + 2 * (k & umask) : ti & amask;    // this group's buffers: grp, grp + 2
                const uint32_t acc_addr = lane_addr + buf * NC;
                const uint32_t full_u32 = smem_u32(&acc_full[buf]), empty_u32 = smem_u32(&acc_empty[buf]);
                const uint32_t fpar = (uint32_t)(k >> ushift) & 1u;
                if (a.stats && a.stats_nstride != 0 && cur_n >= 0 && it.n != cur_n) combine(cur_n);
                cur_n = it.n;
                const int oy = it.ty * TILE_H + my, ox = it.tx * TW + (PAIR ? grp * TILE_W : 0) + mx;
                const bool pvalid = oy < a.OH && ox < a.OW;
                bf16* yp = a.y + ((size_t)(it.n * a.OH + (pvalid ? oy : 0)) * a.OW + (pvalid ? ox : 0)) * a.y_pitch + a.y_coff;
                float* ypf = reinterpret_cast<float*>(a.y) + ((size_t)(it.n * a.OH + (pvalid ? oy : 0)) * a.OW + (pvalid ? ox : 0)) * a.y_pitch + a.y_coff;
                mbar_wait_warp(full_u32, fpar);
                tc_fence_after();
                int dq = d2s_q0, dc = d2s_c0;                 // depth-to-space: quadrant and channel inside it of group g
                for (int g = 0; g < NC / 8; ++g) {
                    const int c = c_begin + g * 8;
                    float v[8];
                    tmem_ld8(acc_addr + g * 8, v);
                    const bool cvalid = c < a.Cout;
                    if (a.d2s_c) {
                        // Conv2DTranspose(2x2, s2): channel group (quadrant dq, channels dc..dc+7) of pixel (oy, ox) is pixel
                        // (2 oy + dq/2, 2 ox + dq%2) of the up-sampled tensor; the C-long bias is added here (no separate pass)
                        const int uy = 2 * oy + (dq >> 1), ux = 2 * ox + (dq & 1);
                        if (pvalid && cvalid && uy < a.d2s_h && ux < a.d2s_w) {
                            if (a.bias) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] += a.bias[dc + i];
                            }
                            Vec8<bf16>::store(a.y + ((size_t)(it.n * a.d2s_h + uy) * a.d2s_w + ux) * a.y_pitch + a.y_coff + dc, v);
                        }
                        dc += 8;
                        if (dc >= a.d2s_c) { dc -= a.d2s_c; ++dq; }
                        continue;
                    }
                    if (a.bias && cvalid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += a.bias[c + i];
                    }
                    if (a.stats) {
                        float w[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { w[i] = pvalid ? v[i] : 0.f; w[8 + i] = w[i] * w[i]; }
                        warp_reduce16(w, lane);
                        if ((lane & 1) == 0) {
                            const int idx = lane >> 1;                       // 0..7 sums, 8..15 squares
                            part[((idx >> 3) * 4 + wq) * NC + g * 8 + (idx & 7)] += w[0];
                        }
                    }
                    if (pvalid && cvalid && !(a.dbg & 4)) {
                        if (a.accumulate) {
                            float o[8];
                            if (a.out_f32) Vec8<float>::load(ypf + c, o); else Vec8<bf16>::load(yp + c, o);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += o[i];
                        }
                        if (a.out_f32) Vec8<float>::store(ypf + c, v); else Vec8<bf16>::store(yp + c, v);
                    }
                }
                tc_fence_before();
                mbar_arrive(empty_u32);
            }
            if (a.stats && cur_n >= 0) combine(a.stats_nstride != 0 ? cur_n : 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc_rt(tmem, (uint32_t)args.tmem_cols);
}

// ---- host side --------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Called by semb_conv2d_fwd_tc (conv_tc.cu) after argument validation, for SEMB_PAD_ZERO geometries.
int conv_tma_launch(const semb_conv_geom* g, const semb_tensor* x, const void* w_packed, const float* bias, const semb_tensor* y,
                    void* stats, int32_t stats_nstride, int32_t stats_cstride, int32_t accumulate, void* stream, int out_f32,
                    int d2s_h, int d2s_w) {
    TmArgs A{};
    TcArgs& a = A.t;
    a.out_f32 = out_f32;
    if (d2s_h > 0) { a.d2s_c = y->C; a.d2s_h = d2s_h; a.d2s_w = d2s_w; }
    a.N = g->N; a.H = g->H; a.W = g->W; a.OH = g->OH; a.OW = g->OW; a.Cin = g->Cin; a.Cout = g->Cout;
    a.R = g->R; a.S = g->S; a.pad_t = g->pad_t; a.pad_l = g->pad_l; a.pad_mode = g->pad_mode;
    a.x = reinterpret_cast<const bf16*>(x->ptr); a.x_pitch = x->pitch; a.x_coff = x->coff;
    a.y = reinterpret_cast<bf16*>(y->ptr); a.y_pitch = y->pitch; a.y_coff = y->coff;
    a.wp = reinterpret_cast<const bf16*>(w_packed); a.bias = bias;
    a.stats = reinterpret_cast<double*>(stats); a.stats_nstride = stats_nstride; a.stats_cstride = stats_cstride;
    a.accumulate = accumulate;
    a.p = tc_plan(g->Cin, g->Cout, g->R * g->S);
    // PAIR mode (see the kernel): every layer whose weights are streamed per K chunk and whose image is wider than one tile
    static const bool pair_on = [] { const char* e = getenv("SEMB_TMA_NO_PAIR"); return !(e && e[0] && e[0] != '0'); }();
    bool pair = pair_on && a.p.kchunks > 1 && a.p.NC <= 128 && g->OW > TILE_W;
    if (pair) {                              // two stages of pair-sized halo planes + weight chunk must fit (tc_plan sizes K chunks for it)
        const size_t pp = (size_t)((TILE_H + g->R - 1) * (2 * TILE_W + g->S - 1) * 16 + 127) / 128 * 128;
        const size_t st = (size_t)(a.p.KC / 8) * pp + (size_t)g->R * g->S * a.p.KC * a.p.NC * 2;
        if (2 * st + (size_t)16 * a.p.NC * sizeof(float) + 128 > 220 * 1024) pair = false;
    }
    const int tw = pair ? 2 * TILE_W : TILE_W;
    a.tiles_x = cdiv(g->OW, tw); a.tiles_y = cdiv(g->OH, TILE_H);
    a.total_tiles = g->N * a.tiles_x * a.tiles_y;
    a.halo_h = TILE_H + g->R - 1; a.halo_w = tw + g->S - 1;
    A.plane_box = a.halo_h * a.halo_w * 16;
    A.plane_pitch = (A.plane_box + 127) / 128 * 128;
    a.plane_bytes = A.plane_pitch;
    const int taps = g->R * g->S;
    a.a_bytes = (a.p.KC / 8) * A.plane_pitch;
    a.b_bytes = taps * a.p.KC * a.p.NC * 2;
    a.b_resident = a.p.kchunks == 1;
    a.stage_bytes = a.a_bytes + (a.b_resident ? 0 : a.b_bytes);
    const size_t fixed = (a.b_resident ? a.b_bytes : 0) + (size_t)16 * a.p.NC * sizeof(float) + 128;
    // ring depth: enough bytes in flight per SM for HBM (>= ~48 KB with two CTAs), within half of the shared memory
    int nst = 2;
    while (nst < TM_MAX_STAGES && (size_t)(nst + 1) * a.stage_bytes + fixed <= 100 * 1024 && (size_t)nst * a.stage_bytes < 48 * 1024) ++nst;
    while (nst > 2 && (size_t)nst * a.stage_bytes + fixed > 220 * 1024) --nst;
    if (pair) {                              // one CTA per SM: as many stages as fit (the weight stream wants bytes in flight)
        nst = 2;
        while (nst < TM_MAX_STAGES && (size_t)(nst + 1) * a.stage_bytes + fixed <= 200 * 1024) ++nst;
    }
    // tuning / ablation knobs, read once per process (scripts/ablate_conv.sh, scripts/sweep_stage.sh)
    struct Knobs { int stages, nacc, dbg, per_sm; };
    static const Knobs knobs = [] {
        auto num = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
        return Knobs{num("SEMB_TMA_STAGES"), num("SEMB_TMA_NACC"), num("SEMB_TC_DEBUG"), num("SEMB_TC_PER_SM")};
    }();
    if (knobs.stages >= 2 && knobs.stages <= TM_MAX_STAGES) nst = knobs.stages;
    A.nmma = (pair || nst >= 4) ? 2 : 1;
    if (A.nmma == 2 && !pair) nst &= ~1;    // two half-rings, one per MMA warp
    A.nstages = nst;
    a.stages = nst;
    A.bsplit = 1;
    if (!a.b_resident) {
        static const int want = [] { const char* e = getenv("SEMB_TMA_BSPLIT"); return e ? atoi(e) : 0; }();
        int q = want < 1 ? (pair ? 4 : 1) : want;
        while (q > 1 && (a.b_bytes % (q * 16)) != 0) --q;
        A.bsplit = q;
    }
    const size_t smem = (size_t)nst * a.stage_bytes + fixed;
    SEMB_REQUIRE(smem <= 220 * 1024, SEMB_EWORKSPACE, "conv_tma: %zu bytes of shared memory needed", smem);
    A.nacc = (a.p.NC <= 64 || pair) ? 4 : 2;
    if (knobs.nacc == 2 || (knobs.nacc == 4 && a.p.NC <= 128)) A.nacc = knobs.nacc;
    const int need = A.nacc * a.p.NC;
    const int ks_fixed = a.p.kchunks == 1 ? (g->Cin + 15) / 16 : 0;
    const int cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
    A.tmem_cols = cols;
    int per_sm = 512 / cols;                                  // TMEM columns
    if ((size_t)per_sm * (smem + 2048) > 220 * 1024) per_sm = (int)(220 * 1024 / (smem + 2048));   // shared memory
    if (per_sm > 2) per_sm = 2;                               // registers: 352 threads x <= 93
    a.dbg = knobs.dbg;
    if (knobs.per_sm >= 1 && knobs.per_sm < per_sm) per_sm = knobs.per_sm;
    if (per_sm < 1) per_sm = 1;
    int gx = 148 * per_sm;
    if (gx > a.total_tiles) gx = a.total_tiles;
    dim3 grid(gx, a.p.nchunks);

    EncodeTiledFn enc = encode_tiled();
    SEMB_REQUIRE(enc != nullptr, SEMB_ECUDA, "conv_tma: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap xmap;
    const cuuint64_t dims[4] = {(cuuint64_t)x->pitch, (cuuint64_t)g->W, (cuuint64_t)g->H, (cuuint64_t)g->N};
    const cuuint64_t strides[3] = {(cuuint64_t)x->pitch * 2, (cuuint64_t)g->W * x->pitch * 2, (cuuint64_t)g->H * g->W * x->pitch * 2};
    const cuuint32_t box[4] = {8u, (cuuint32_t)a.halo_w, (cuuint32_t)a.halo_h, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult cr = enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x->ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SEMB_REQUIRE(cr == CUDA_SUCCESS, SEMB_ECUDA, "conv_tma: cuTensorMapEncodeTiled failed (%d) for pitch %d W %d H %d N %d", (int)cr,
                 x->pitch, g->W, g->H, g->N);

    cudaError_t e = cudaSuccess;
#define SEMB_TM_LAUNCH3(NCT, KR, KS)                                                                                     \
    if (colsplit) {                                                                                                      \
        e = cudaFuncSetAttribute(conv_tma_kernel<16, KR, KS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) e = launch_pdl(conv_tma_kernel<16, KR, KS, true>, grid, dim3(TM_THREADS), smem, as_stream(stream), A, xmap); \
    } else {                                                                                                             \
        e = cudaFuncSetAttribute(conv_tma_kernel<NCT, KR, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (e == cudaSuccess) e = launch_pdl(conv_tma_kernel<NCT, KR, KS>, grid, dim3(TM_THREADS), smem, as_stream(stream), A, xmap); \
    }
#define SEMB_TM_LAUNCH2(NCT, KR)                                                                                         \
    switch (ks_fixed) {                                                                                                  \
        case 1: SEMB_TM_LAUNCH3(NCT, KR, 1) break;                                                                       \
        case 2: SEMB_TM_LAUNCH3(NCT, KR, 2) break;                                                                       \
        case 3: SEMB_TM_LAUNCH3(NCT, KR, 3) break;                                                                       \
        case 4: SEMB_TM_LAUNCH3(NCT, KR, 4) break;                                                                       \
        default: SEMB_TM_LAUNCH3(NCT, KR, 0) break;                                                                      \
    }
#define SEMB_TM_LAUNCH(NCT)                                                                                              \
    if (g->R == 3) { SEMB_TM_LAUNCH2(NCT, 3) } else { SEMB_TM_LAUNCH2(NCT, 1) }
    // NC == 32 in one group would need 64 moment registers per thread (spills under the 92-register cap): the two epilogue
    // groups split the columns instead (CS); wider layers take the generic epilogue
    static const bool cs_on = [] { const char* e = getenv("SEMB_TMA_NO_COLSPLIT"); return !(e && e[0] && e[0] != '0'); }();
    const bool colsplit = cs_on && a.p.NC == 32 && A.nacc == 4 && !pair && !a.d2s_c;
    if (pair) {
        if (g->R == 3) {
            e = cudaFuncSetAttribute(conv_tma_kernel<0, 3, 0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = launch_pdl(conv_tma_kernel<0, 3, 0, false, true>, grid, dim3(TM_THREADS), smem, as_stream(stream), A, xmap);
        } else {
            e = cudaFuncSetAttribute(conv_tma_kernel<0, 1, 0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = launch_pdl(conv_tma_kernel<0, 1, 0, false, true>, grid, dim3(TM_THREADS), smem, as_stream(stream), A, xmap);
        }
    } else if ((a.p.NC == 16 && !a.d2s_c) || colsplit) { SEMB_TM_LAUNCH(16) } else { SEMB_TM_LAUNCH(0) }
#undef SEMB_TM_LAUNCH
#undef SEMB_TM_LAUNCH2
#undef SEMB_TM_LAUNCH3
    if (e != cudaSuccess) { set_error("conv_tma: launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return SEMB_ECUDA; }
    return check_launch("conv_tma");
}

}  // namespace semb
