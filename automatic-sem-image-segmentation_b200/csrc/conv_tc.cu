// Implicit-GEMM convolutions on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM).
//
// im2col-free: a CTA stages the NHWC halo of its 16x8 output-pixel tile ONCE into shared memory in
// "channel-group planes" [Cin/8][halo_h][halo_w][8 x bf16], which is exactly the no-swizzle K-major UMMA
// canonical layout (8 consecutive pixels of one row = one 128-byte core matrix, SBO = halo row pitch,
// LBO = plane pitch).  The 3x3 taps are then nine shared-memory DESCRIPTORS that point at shifted starts
// inside the same halo tile -- no im2col matrix, no re-load per tap.  Weights are pre-packed on the device
// into the matching [tap][Cin/8][Cout][8] image.  One elected thread issues the MMAs; all four warps drain
// the 128-lane accumulator with tcgen05.ld in the epilogue (bias, fp32 BatchNorm/InstanceNorm moments,
// bf16 store at a channel offset of the destination buffer = skip-concat for free).
//
// Replaces F.conv2d / its data gradient under keras.layers.Conv2D for the stride-1 3x3 and 1x1 layers
// (UNet_Segmentation.py:421,465-468,490-499; CycleGAN.py:327,333).
#include "tc_common.cuh"
#include <stdlib.h>
#include <math.h>

namespace semb {


// ---- weight packing ----------------------------------------------------------------------------------
// dst[nchunk][kchunk][tap][KC/8][NC][8] (bf16)  <-  w[r][s][ci][co] (fp32 HWIO), zero padded.
// flip: the packed operator is the stride-1 data gradient: k runs over co, n over ci, taps mirrored.
__global__ void pack_weights_kernel(const float* __restrict__ w, int R, int S, int Cin, int Cout, int flip, TcPlan p,
                                    bf16* __restrict__ dst, long long total) {
    const int taps = R * S;
    const int K = flip ? Cout : Cin;      // reduction channels of the packed operator
    const int N = flip ? Cin : Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int ke = (int)(t % 8); t /= 8;
        const int nl = (int)(t % p.NC); t /= p.NC;
        const int k8 = (int)(t % (p.KC / 8)); t /= (p.KC / 8);
        const int tap = (int)(t % taps); t /= taps;
        const int kch = (int)(t % p.kchunks); t /= p.kchunks;
        const int nch = (int)t;
        const int k = kch * p.KC + k8 * 8 + ke, n = nch * p.NC + nl;
        float v = 0.f;
        if (k < K && n < N) {
            if (!flip) v = w[((size_t)tap * Cin + k) * Cout + n];
            else v = w[((size_t)(taps - 1 - tap) * Cin + n) * Cout + k];
        }
        dst[i] = __float2bfloat16_rn(v);
    }
}


// All weight images of a network in ONE launch (a train step re-packs ~120 tensors after Adam): job table in device memory.
struct PackJob {
    const float* w;
    bf16* dst;
    int R, S, Cin, Cout, flip;
    TcPlan p;
    long long total;
    int block0, nblocks;
};

__global__ void pack_weights_batch_kernel(const PackJob* __restrict__ jobs, int njobs) {
    pdl_trigger();
    pdl_wait();
    int j = 0;
    while (j + 1 < njobs && (int)blockIdx.x >= jobs[j + 1].block0) ++j;
    const PackJob job = jobs[j];
    const int taps = job.R * job.S;
    const int K = job.flip ? job.Cout : job.Cin;
    const int N = job.flip ? job.Cin : job.Cout;
    const TcPlan p = job.p;
    const long long stride = (long long)job.nblocks * blockDim.x;
    for (long long i = (long long)(blockIdx.x - job.block0) * blockDim.x + threadIdx.x; i < job.total; i += stride) {
        long long t = i;
        const int ke = (int)(t % 8); t /= 8;
        const int nl = (int)(t % p.NC); t /= p.NC;
        const int k8 = (int)(t % (p.KC / 8)); t /= (p.KC / 8);
        const int tap = (int)(t % taps); t /= taps;
        const int kch = (int)(t % p.kchunks); t /= p.kchunks;
        const int nch = (int)t;
        const int k = kch * p.KC + k8 * 8 + ke, n = nch * p.NC + nl;
        float v = 0.f;
        if (k < K && n < N) {
            if (!job.flip) v = job.w[((size_t)tap * job.Cin + k) * job.Cout + n];
            else v = job.w[((size_t)(taps - 1 - tap) * job.Cin + n) * job.Cout + k];
        }
        job.dst[i] = __float2bfloat16_rn(v);
    }
}

// ---- forward / data-gradient conv: persistent, warp-specialised -----------------------------------------------
//   warps 0-3  epilogue   : drain TMEM (lane = output pixel), bias, moments, bf16 store at the channel offset
//   warp  4    MMA issuer : one elected thread, tcgen05.mma per (tap, 16-channel step), tcgen05.commit
//   warps 5-8  producers  : cp.async the halo tile of the next (tile, channel-chunk) into a ring of stages
// The accumulator is double-buffered in TMEM (2 x NC columns) so the epilogue of tile i overlaps the MMAs of tile
// i+1 and the loads of tile i+2.  When all input channels fit one chunk the packed weights are loaded ONCE per CTA.
constexpr int FW_THREADS = 288;
constexpr int FW_PRODUCER0 = 160;      // first producer thread

// KR: kernel size (1 or 3), compile-time so that the MMA issue loop unrolls into straight-line code.
template <int COLS, int STAGES, int NCT, int KR>
__global__ void __launch_bounds__(FW_THREADS, NCT == 32 ? 2 : 3) conv_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2], b_full;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = a.p.KC, NC = NCT > 0 ? NCT : a.p.NC, kchunks = a.p.kchunks;
    const int nchunk = blockIdx.y;
    const int halo_pix = a.halo_h * a.halo_w;
    // smem: [resident B (optional)] [STAGES x (A chunk [+ B chunk])] [moment partials 4 x 2 x NC floats]
    uint8_t* ring = smem + (a.b_resident ? a.b_bytes : 0);
    float* part = reinterpret_cast<float*>(ring + (size_t)STAGES * a.stage_bytes);
    const int ntiles = (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 4) tmem_alloc<COLS>(smem_u32(&tmem_slot));
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&full_bar[i]), 128); mbar_init(smem_u32(&empty_bar[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 128); }
        mbar_init(smem_u32(&b_full), 128);
    }
    for (int i = tid; i < 8 * NC; i += FW_THREADS) part[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t ring_u32 = smem_u32(ring);

    if (warp >= 5) {
        // ============================== producers ==============================
        const int ptid = tid - FW_PRODUCER0;
        if (a.b_resident) {
            const uint4* src = reinterpret_cast<const uint4*>(a.wp + (size_t)nchunk * kchunks * a.R * a.S * KC * NC);
            const uint32_t dst = smem_u32(smem);
            for (int idx = ptid; idx < a.b_bytes / 16; idx += 128) cp_async16(dst + idx * 16, src + idx, true);
            cp_async_commit();
            cp_async_wait<0>();
            fence_proxy_async();
            mbar_arrive(smem_u32(&b_full));
        }
        // a thread owns (at most) two fixed halo pixels; only the tile origin changes from tile to tile
        int hy[2], hx[2];
        bool own[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int pix = ptid + 128 * j;
            own[j] = pix < halo_pix;
            hy[j] = pix / a.halo_w;
            hx[j] = pix - hy[j] * a.halo_w;
        }
        TileIter it(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
        int stage = 0;
        uint32_t phase = 1;                      // waits on empty_bar start with the "previous phase" parity
        for (int ti = 0; ti < ntiles; ++ti, it.next()) {
            const int y0 = it.ty * TILE_H - a.pad_t, x0 = it.tx * TILE_W - a.pad_l;
            const bf16* src[2];
            bool v[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int iy = y0 + hy[j], ix = x0 + hx[j];
                if (a.pad_mode == SEMB_PAD_REFLECT) {
                    if (iy > -a.H && iy < 2 * a.H - 1) iy = reflect_index(iy, a.H);
                    if (ix > -a.W && ix < 2 * a.W - 1) ix = reflect_index(ix, a.W);
                }
                v[j] = own[j] && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
                src[j] = a.x + ((size_t)(it.n * a.H + (v[j] ? iy : 0)) * a.W + (v[j] ? ix : 0)) * a.x_pitch + a.x_coff;
            }
            for (int kc = 0; kc < kchunks; ++kc) {
                mbar_wait_warp(smem_u32(&empty_bar[stage]), phase);
                const int c0 = kc * KC;
                const int planes = 2 * ((min(KC, a.Cin - c0) + 15) / 16);
                const int real = min(planes, (a.Cin - c0) >> 3);       // planes beyond Cin are zero-filled
                const uint32_t sbase = ring_u32 + stage * a.stage_bytes;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (own[j] && !(a.dbg & 1)) {
                        const uint32_t dst = sbase + (ptid + 128 * j) * 16;
                        const bf16* s = src[j] + c0;
                        for (int k8 = 0; k8 < planes; ++k8) cp_async16(dst + k8 * a.plane_bytes, s + k8 * 8, v[j] && k8 < real);
                    }
                }
                if (!a.b_resident) {
                    const uint4* bsrc = reinterpret_cast<const uint4*>(a.wp + ((size_t)nchunk * kchunks + kc) * a.R * a.S * KC * NC);
                    const uint32_t dst = sbase + a.a_bytes;
                    for (int idx = ptid; idx < a.b_bytes / 16; idx += 128) cp_async16(dst + idx * 16, bsrc + idx, true);
                }
                cp_async_arrive_noinc(smem_u32(&full_bar[stage]));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        cp_async_wait<0>();
    } else if (warp == 4) {
        // ============================== MMA issuer (whole warp converged, elect.sync per instruction) ===============
        {
            // The single issuing thread is the critical path of a tile (ncu, round 1: ~50 instructions per MMA with
            // run-time tap loops): taps and K steps are unrolled, the descriptors differ only in their low words.
            constexpr int HALO_W = TILE_W + KR - 1;
            const uint32_t idesc = instr_desc(128, NC, 0, 0);
            const uint64_t ad0 = smem_desc(ring_u32, a.plane_bytes, HALO_W * 16);
            const uint64_t bd0 = smem_desc(a.b_resident ? smem_u32(smem) : ring_u32 + a.a_bytes, NC * 16, 128);
            const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
            const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
            const uint32_t pl2 = 2 * ((uint32_t)a.plane_bytes >> 4);  // 16-byte units per K step of the A planes
            const uint32_t nc2 = 2 * (uint32_t)NC;                    // ... of the weight image
            const uint32_t stage16 = (uint32_t)a.stage_bytes >> 4;
            const uint32_t btap = (uint32_t)(KC / 8) * NC;            // 16-byte units between taps of the weight image
            if (a.b_resident) mbar_wait(smem_u32(&b_full), 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int ti = 0; ti < ntiles; ++ti) {
                const int buf = ti & 1;
                mbar_wait(smem_u32(&acc_empty[buf]), ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem + buf * NC;
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase);
                    fence_proxy_async();          // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
                    tc_fence_after();
                    const int ksteps = (min(KC, a.Cin - kc * KC) + 15) / 16;
                    const uint32_t soff = (uint32_t)stage * stage16;
                    const uint32_t a_lo = a_lo0 + soff;
                    const uint32_t b_lo = b_lo0 + (a.b_resident ? 0u : soff);
#pragma unroll
                    for (int tap = 0; tap < KR * KR; ++tap) {
                        const uint32_t at = a_lo + (uint32_t)((tap / KR) * HALO_W + (tap % KR));
                        const uint32_t bt = b_lo + (uint32_t)tap * btap;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {            // KC <= 64: at most four K steps per chunk
                            if (ks < ksteps && !(a.dbg & 2))
                                umma_bf16_elect(dcol, at + ks * pl2, a_hi, bt + ks * nc2, b_hi, idesc, (tap | ks) != 0 ? 1u : (uint32_t)(kc != 0));
                        }
                    }
                    umma_commit_elect(smem_u32(&empty_bar[stage]));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_elect(smem_u32(&acc_full[buf]));
            }
        }
    } else {
        // ============================== epilogue ==============================
        const int c_begin = nchunk * NC;
        int cur_n = -1;
        auto combine = [&](int n) {
            // 128 epilogue threads: combine the four warps' partial moments in a fixed order, one fp64 atomic per channel
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = tid; i < NC; i += 128) {
                const int c = c_begin + i;
                if (c < a.Cout) {
                    const float t1 = ((part[i] + part[NC + i]) + part[2 * NC + i]) + part[3 * NC + i];
                    const float t2 = ((part[4 * NC + i] + part[5 * NC + i]) + part[6 * NC + i]) + part[7 * NC + i];
                    double* st = a.stats + (size_t)n * a.stats_nstride + c;
                    atomicAdd(st, (double)t1);
                    atomicAdd(st + a.stats_cstride, (double)t2);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = tid; i < 8 * NC; i += 128) part[i] = 0.f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        };
        const int m = warp * 32 + lane;
        const int my = m / TILE_W, mx = m % TILE_W;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        TileIter it(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
        if constexpr (NCT > 0) {
            float s1[NCT], s2[NCT];
#pragma unroll
            for (int i = 0; i < NCT; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
            auto flush_regs = [&](int n) {
                // lane totals -> warp totals (halving butterfly, fixed order), even lanes publish one value each
#pragma unroll
                for (int q = 0; q < NCT / 16; ++q) {
                    float w[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = s1[q * 16 + i];
                    warp_reduce16(w, lane);
                    if ((lane & 1) == 0) part[warp * NC + q * 16 + (lane >> 1)] = w[0];
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = s2[q * 16 + i];
                    warp_reduce16(w, lane);
                    if ((lane & 1) == 0) part[(4 + warp) * NC + q * 16 + (lane >> 1)] = w[0];
                }
#pragma unroll
                for (int i = 0; i < NCT; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
                combine(n);
            };
            for (int ti = 0; ti < ntiles; ++ti, it.next()) {
                const int buf = ti & 1;
                if (a.stats && a.stats_nstride != 0 && cur_n >= 0 && it.n != cur_n) flush_regs(cur_n);
                cur_n = it.n;
                const int oy = it.ty * TILE_H + my, ox = it.tx * TILE_W + mx;
                const bool pvalid = oy < a.OH && ox < a.OW;
                bf16* yp = a.y + ((size_t)(it.n * a.OH + (pvalid ? oy : 0)) * a.OW + (pvalid ? ox : 0)) * a.y_pitch + a.y_coff;
                mbar_wait_warp(smem_u32(&acc_full[buf]), (ti >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < NCT / 16; ++h) {
                    float v16[16];
                    tmem_ld16(lane_base + buf * NC + h * 16, v16);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int c = h * 16 + half * 8;
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
                            for (int i = 0; i < 8; ++i) { s1[c + i] += v[i]; s2[c + i] = fmaf(v[i], v[i], s2[c + i]); }
                        }
                        if (pvalid && cvalid && !(a.dbg & 4)) {
                            if (a.accumulate) {
                                float o[8];
                                Vec8<bf16>::load(yp + c, o);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] += o[i];
                            }
                            Vec8<bf16>::store(yp + c, v);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(smem_u32(&acc_empty[buf]));
            }
            if (a.stats && cur_n >= 0) flush_regs(a.stats_nstride != 0 ? cur_n : 0);
        } else {
            for (int ti = 0; ti < ntiles; ++ti, it.next()) {
                const int buf = ti & 1;
                if (a.stats && a.stats_nstride != 0 && cur_n >= 0 && it.n != cur_n) combine(cur_n);
                cur_n = it.n;
                const int oy = it.ty * TILE_H + my, ox = it.tx * TILE_W + mx;
                const bool pvalid = oy < a.OH && ox < a.OW;
                bf16* yp = a.y + ((size_t)(it.n * a.OH + (pvalid ? oy : 0)) * a.OW + (pvalid ? ox : 0)) * a.y_pitch + a.y_coff;
                mbar_wait_warp(smem_u32(&acc_full[buf]), (ti >> 1) & 1);
                tc_fence_after();
                for (int g = 0; g < NC / 8; ++g) {
                    const int c = c_begin + g * 8;
                    float v[8];
                    tmem_ld8(lane_base + buf * NC + g * 8, v);
                    const bool cvalid = c < a.Cout;
                    if (a.bias && cvalid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += a.bias[c + i];
                    }
                    if (a.stats) {
                        // 16 per-pixel values (8 sums, 8 squares) -> warp totals with a halving butterfly: 8+4+2+1+1 = 16
                        // shuffles instead of 80; even lane l ends up with the total of value (l >> 1).
                        float w[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { w[i] = pvalid ? v[i] : 0.f; w[8 + i] = w[i] * w[i]; }
                        warp_reduce16(w, lane);
                        if ((lane & 1) == 0) {
                            const int idx = lane >> 1;                       // 0..7 sums, 8..15 squares
                            part[((idx >> 3) * 4 + warp) * NC + g * 8 + (idx & 7)] += w[0];
                        }
                    }
                    if (pvalid && cvalid) {
                        if (a.accumulate) {
                            float o[8];
                            Vec8<bf16>::load(yp + c, o);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += o[i];
                        }
                        Vec8<bf16>::store(yp + c, v);
                    }
                }
                tc_fence_before();
                mbar_arrive(smem_u32(&acc_empty[buf]));
            }
            if (a.stats && cur_n >= 0) combine(a.stats_nstride != 0 ? cur_n : 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<COLS>(tmem);
}

// ---- weight gradient on the tensor cores ------------------------------------------------------------------
// dW[tap][ci][co] = sum_pixels x[pixel+tap][ci] * dy[pixel][co]  as  D[M=co][N=ci] += A[co][k] * B[ci][k], k = pixels.
// Both operands are read straight from their NHWC channel-group planes as MN-MAJOR UMMA operands (the same
// shared-memory image the forward kernel uses as K-major): 8 consecutive pixels of a row are the 8 K-rows of a
// core matrix, LBO = next pixel row, SBO = next channel-group plane.  Each tap owns a block of TMEM columns.
//
// Persistent, warp-specialised: 4 producer warps stream (dy tile, x halo tile) pairs through a STAGES-deep
// cp.async ring; one elected thread of a 5th warp issues the tcgen05.mma's (K = 16 pixels each) and releases
// ring slots with tcgen05.commit; the accumulator stays in TMEM over ALL pixel tiles of the CTA (split-K over
// pixels) and is flushed once with coalesced fp32 reductions into the HWIO gradient.
constexpr int WG_PRODUCERS = 128;
constexpr int WG_THREADS = WG_PRODUCERS + 32;

struct WgPlan { int NCH, nchunks, tpp, tapgroups, mtiles, cols, splits, stage_bytes, planes_a_max, stages, smem_bytes; };

static inline WgPlan wg_plan(int Cin, int Cout, int taps, long long total_tiles, int plane_a, int plane_b) {
    WgPlan p;
    const int cin16 = (Cin + 15) / 16 * 16;
    p.NCH = cin16 < 128 ? cin16 : 128;          // keeps a stage (dy tile + x halo) below 80 KB
    p.nchunks = (cin16 + p.NCH - 1) / p.NCH;
    int tpp = 512 / p.NCH;
    if (tpp > taps) tpp = taps;
    p.tpp = tpp;
    p.tapgroups = (taps + p.tpp - 1) / p.tpp;
    p.mtiles = (Cout + 127) / 128;
    const int c = p.tpp * p.NCH;
    p.cols = c <= 32 ? 32 : (c <= 64 ? 64 : (c <= 128 ? 128 : (c <= 256 ? 256 : 512)));
    p.planes_a_max = Cout >= 128 ? 16 : (Cout + 7) / 8;
    p.stage_bytes = p.planes_a_max * plane_a + (p.NCH / 8) * plane_b;
    // the M=128 MMA always reads 16 A planes; planes beyond Cout are garbage rows that are never flushed, but the
    // reads must stay inside the allocation
    const int slack = 16 * plane_a > p.stage_bytes ? 16 * plane_a - p.stage_bytes : 0;
    p.stages = (size_t)p.stage_bytes * 3 + slack <= 200 * 1024 ? 3 : 2;
    p.smem_bytes = p.stage_bytes * p.stages + slack;
    const int work = p.mtiles * p.tapgroups * p.nchunks;
    const int per_sm = (p.cols <= 256 && p.smem_bytes <= 100 * 1024) ? 2 : 1;
    long long s = (148LL * per_sm) / work;
    if (s > total_tiles) s = total_tiles;
    if (s < 1) s = 1;
    p.splits = (int)s;
    return p;
}

struct WgArgs {
    int N, H, W, OH, OW, Cin, Cout, R, S, pad_t, pad_l, pad_mode;
    const bf16* x; int x_pitch, x_coff;
    const bf16* dy; int dy_pitch, dy_coff;
    float* dw;
    int tiles_x, tiles_y, total_tiles;
    WgPlan p;
    int plane_a, plane_b, halo_h, halo_w;
};


template <int COLS, int WG_STAGES>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_kernel(const WgArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[WG_STAGES], empty_bar[WG_STAGES], done_bar;
    __shared__ uint32_t tmem_slot;

    pdl_trigger();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int taps = a.R * a.S;
    const int NCH = a.p.NCH;
    int wi = blockIdx.y;
    const int nc = wi % a.p.nchunks; wi /= a.p.nchunks;
    const int tg = wi % a.p.tapgroups; wi /= a.p.tapgroups;
    const int mt = wi;
    const int tap0 = tg * a.p.tpp;
    const int ntaps = min(a.p.tpp, taps - tap0);
    const int co0 = mt * 128, ci0 = nc * NCH;
    const int planes_a = min(16, (a.Cout - co0 + 7) / 8);
    const int planes_b = min(NCH / 8, (a.Cin - ci0 + 7) / 8);
    const int halo_pix = a.halo_h * a.halo_w;
    const int b_off = a.p.planes_a_max * a.plane_a;                 // B planes follow the A planes inside a stage
    const int ntiles = (a.total_tiles - (int)blockIdx.x + a.p.splits - 1) / a.p.splits;   // tiles of this CTA

    if (warp == 4) tmem_alloc<COLS>(smem_u32(&tmem_slot));
    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(smem_u32(&full_bar[i]), WG_PRODUCERS); mbar_init(smem_u32(&empty_bar[i]), 1); }
        mbar_init(smem_u32(&done_bar), 1);
    }
    // N-side planes beyond Cin are never loaded: keep them finite (they only feed columns that are never flushed)
    for (int st = 0; st < WG_STAGES; ++st)
        for (int idx = tid; idx < (NCH / 8 - planes_b) * halo_pix; idx += WG_THREADS) {
            const int k8 = planes_b + idx / halo_pix, pix = idx % halo_pix;
            *reinterpret_cast<uint4*>(smem + (size_t)st * a.p.stage_bytes + b_off + (size_t)k8 * a.plane_b + pix * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t smem_base = smem_u32(smem);
    pdl_wait();

    if (warp < 4) {
        // ===================== producers =====================
        for (int it = 0; it < ntiles; ++it) {
            const int stage = it % WG_STAGES;
            mbar_wait_warp(smem_u32(&empty_bar[stage]), ((it / WG_STAGES) & 1) ^ 1);
            const int t = blockIdx.x + it * a.p.splits;
            const int n = t / (a.tiles_x * a.tiles_y);
            const int trem = t % (a.tiles_x * a.tiles_y);
            const int y0 = (trem / a.tiles_x) * TILE_H, x0 = (trem % a.tiles_x) * TILE_W;
            const uint32_t sbase = smem_base + stage * a.p.stage_bytes;
            {   // A: dy tile, one pixel per producer thread, all of its channel-group planes
                const int oy = y0 + tid / TILE_W, ox = x0 + tid % TILE_W;
                const bool v = oy < a.OH && ox < a.OW;
                const bf16* src = a.dy + ((size_t)(n * a.OH + (v ? oy : 0)) * a.OW + (v ? ox : 0)) * a.dy_pitch + a.dy_coff + co0;
                const uint32_t dst = sbase + tid * 16;
                for (int k8 = 0; k8 < planes_a; ++k8) cp_async16(dst + k8 * a.plane_a, src + k8 * 8, v);
            }
            for (int pix = tid; pix < halo_pix; pix += WG_PRODUCERS) {   // B: x halo tile
                const int hy = pix / a.halo_w, hx = pix % a.halo_w;
                int iy = y0 - a.pad_t + hy, ix = x0 - a.pad_l + hx;
                if (a.pad_mode == SEMB_PAD_REFLECT) {
                    if (iy > -a.H && iy < 2 * a.H - 1) iy = reflect_index(iy, a.H);
                    if (ix > -a.W && ix < 2 * a.W - 1) ix = reflect_index(ix, a.W);
                }
                const bool v = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
                const bf16* src = a.x + ((size_t)(n * a.H + (v ? iy : 0)) * a.W + (v ? ix : 0)) * a.x_pitch + a.x_coff + ci0;
                const uint32_t dst = sbase + b_off + pix * 16;
                for (int k8 = 0; k8 < planes_b; ++k8) cp_async16(dst + k8 * a.plane_b, src + k8 * 8, v);
            }
            cp_async_arrive_noinc(smem_u32(&full_bar[stage]));
        }
        cp_async_wait<0>();
    } else {
        // ===================== MMA issuer (whole warp converged, elect.sync per instruction) =====================
        const uint32_t idesc = instr_desc(128, NCH, 1, 1);
        const uint64_t ad0 = smem_desc(smem_base, TILE_W * 16, a.plane_a);
        const uint64_t bd0 = smem_desc(smem_base + b_off, a.halo_w * 16, a.plane_b);
        const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
        const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
        for (int it = 0; it < ntiles; ++it) {
            const int stage = it % WG_STAGES;
            mbar_wait(smem_u32(&full_bar[stage]), (it / WG_STAGES) & 1);
            fence_proxy_async();
            tc_fence_after();
            const uint32_t soff = (uint32_t)(stage * a.p.stage_bytes) >> 4;
            for (int tl = 0; tl < ntaps; ++tl) {
                const int tap = tap0 + tl, r = tap / a.S, s = tap % a.S;
#pragma unroll
                for (int ks = 0; ks < TILE_H / 2; ++ks) {       // 16 pixels (two rows of 8) per MMA
                    umma_bf16_elect(tmem + tl * NCH, a_lo0 + soff + ks * (2 * TILE_W), a_hi,
                                    b_lo0 + soff + ((2 * ks + r) * a.halo_w + s), b_hi, idesc, (it | ks) != 0);
                }
            }
            umma_commit_elect(smem_u32(&empty_bar[stage]));
        }
        umma_commit_elect(smem_u32(&done_bar));
    }
    // ---- flush: TMEM lane = co, columns = [tap][ci]; lanes of a warp hit consecutive co -> coalesced reductions
    if (warp < 4) {
        mbar_wait_warp(smem_u32(&done_bar), 0);
        tc_fence_after();
        if (ntiles > 0) {
            const int co = co0 + warp * 32 + lane;
            for (int tl = 0; tl < ntaps; ++tl) {
                const int tap = tap0 + tl;
                for (int g = 0; g < NCH / 8; ++g) {
                    float v[8];
                    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + tl * NCH + g * 8, v);
                    if (co < a.Cout) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int ci = ci0 + g * 8 + i;
                            if (ci < a.Cin) atomicAdd(a.dw + ((size_t)tap * a.Cin + ci) * a.Cout + co, v[i]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<COLS>(tmem);
}


// ---- weight gradient of the 3x3 layers with few input channels: horizontal taps stacked along M -------------------
// With Cin <= 40 the per-tap formulation above issues 72 tiny MMAs per pixel tile, each of which re-reads the 128-row
// dy operand from shared memory: the tensor pipe is busy 90 % of the time moving garbage rows (ncu, round 1).  Here the
// M side of the MMA is the x halo instead, replicated three times in shared memory -- plane (s, k8) holds rows
// [y0-1, y0+17) x columns [x0-1+s, x0+7+s) of channel group k8 -- so that ONE descriptor with a uniform 8-row-group
// stride covers the three horizontal taps of all channel groups (M = 3*Cin <= 120 rows), the vertical tap r is a
// 128-byte shift of the descriptor start, and dy (N = Cout) is read once per K step: 24 MMAs per tile instead of 72,
// D[r] = [(s, ci)][co] in three TMEM column blocks.  Producers / ring / split-K / flush as in wgrad_tc_kernel.
constexpr int WS_ROWS = TILE_H + 2;
constexpr int WS_PLANE_A = WS_ROWS * TILE_W * 16;      // 2304 bytes: [18 rows][8 pixels][8 channels]
constexpr int WS_PLANE_B = TILE_H * TILE_W * 16;       // 2048 bytes: [16 rows][8 pixels][8 channels]

struct WsArgs {
    int N, H, W, OH, OW, Cin, Cout, pad_t, pad_l, pad_mode;
    const bf16* x; int x_pitch, x_coff;
    const bf16* dy; int dy_pitch, dy_coff;
    float* dw;
    int tiles_x, tiles_y, total_tiles;
    int p, NB, a_bytes, stage_bytes, splits;
};

template <int COLS, int WG_STAGES>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_stacked_kernel(const WsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[WG_STAGES], empty_bar[WG_STAGES], done_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = a.p, NB = a.NB;
    const int ntiles = (a.total_tiles - (int)blockIdx.x + a.splits - 1) / a.splits;

    if (warp == 4) tmem_alloc<COLS>(smem_u32(&tmem_slot));
    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(smem_u32(&full_bar[i]), WG_PRODUCERS); mbar_init(smem_u32(&empty_bar[i]), 1); }
        mbar_init(smem_u32(&done_bar), 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < 4) {
        // ===================== producers =====================
        // x: a thread owns halo slots q = tid and (tid < 16) q = tid + 128 of the 18 x 8 grid, for each of the 3 shifts
        const int qrow[2] = {tid >> 3, (tid + 128) >> 3};
        const int qcol = tid & 7;
        const int nslots = tid < WS_ROWS * TILE_W - 128 ? 2 : 1;
        TileIter it(blockIdx.x, a.splits, a.tiles_x, a.tiles_y);
        int stage = 0;
        uint32_t phase = 1;
        for (int i = 0; i < ntiles; ++i, it.next()) {
            mbar_wait_warp(smem_u32(&empty_bar[stage]), phase);
            const int y0 = it.ty * TILE_H, x0 = it.tx * TILE_W;
            const uint32_t sbase = smem_base + stage * a.stage_bytes;
            {   // dy tile: one pixel per thread, all of its channel groups
                const int oy = y0 + (tid >> 3), ox = x0 + qcol;
                const bool v = oy < a.OH && ox < a.OW;
                const bf16* src = a.dy + ((size_t)(it.n * a.OH + (v ? oy : 0)) * a.OW + (v ? ox : 0)) * a.dy_pitch + a.dy_coff;
                const uint32_t dst = sbase + a.a_bytes + tid * 16;
                for (int k8 = 0; k8 < (a.Cout >> 3); ++k8) cp_async16(dst + k8 * WS_PLANE_B, src + k8 * 8, v);
            }
            for (int j = 0; j < nslots; ++j) {
                int iy = y0 - a.pad_t + qrow[j];
                const bool rowref = a.pad_mode == SEMB_PAD_REFLECT;
                if (rowref && iy > -a.H && iy < 2 * a.H - 1) iy = reflect_index(iy, a.H);
                const bool vy = iy >= 0 && iy < a.H;
                const bf16* rowp = a.x + (size_t)(it.n * a.H + (vy ? iy : 0)) * a.W * a.x_pitch + a.x_coff;
                const uint32_t dst = sbase + (tid + 128 * j) * 16;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    int ix = x0 - a.pad_l + qcol + s;
                    if (rowref && ix > -a.W && ix < 2 * a.W - 1) ix = reflect_index(ix, a.W);
                    const bool v = vy && ix >= 0 && ix < a.W;
                    const bf16* src = rowp + (size_t)(v ? ix : 0) * a.x_pitch;
                    const uint32_t d = dst + s * p * WS_PLANE_A;
                    for (int k8 = 0; k8 < p; ++k8) cp_async16(d + k8 * WS_PLANE_A, src + k8 * 8, v);
                }
            }
            cp_async_arrive_noinc(smem_u32(&full_bar[stage]));
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        cp_async_wait<0>();
    } else {
        // ===================== MMA issuer (whole warp converged, elect.sync per instruction) =====================
        const uint32_t idesc = instr_desc(128, NB, 1, 1);
        const uint64_t ad0 = smem_desc(smem_base, TILE_W * 16, WS_PLANE_A);
        const uint64_t bd0 = smem_desc(smem_base + a.a_bytes, TILE_W * 16, WS_PLANE_B);
        const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
        const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
        const uint32_t stage16 = (uint32_t)a.stage_bytes >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < ntiles; ++i) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            fence_proxy_async();
            tc_fence_after();
            const uint32_t soff = (uint32_t)stage * stage16;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int ks = 0; ks < TILE_H / 2; ++ks) {           // 16 pixels (two rows of 8) per MMA
                    umma_bf16_elect(tmem + r * NB, a_lo0 + soff + (uint32_t)(2 * ks + r) * TILE_W, a_hi,
                                    b_lo0 + soff + (uint32_t)(2 * ks) * TILE_W, b_hi, idesc, (i | ks) != 0);
                }
            }
            umma_commit_elect(smem_u32(&empty_bar[stage]));
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_elect(smem_u32(&done_bar));
    }
    // ---- flush: TMEM lane = (s * p + k8) * 8 + c, columns = [r][co]; a thread owns one (s, ci) row of the gradient
    if (warp < 4) {
        mbar_wait_warp(smem_u32(&done_bar), 0);
        tc_fence_after();
        if (ntiles > 0) {
            const int l = warp * 32 + lane;
            const int g = l >> 3;
            const bool rvalid = g < 3 * p;
            const int s = g / p, k8 = g - s * p;
            const int ci = k8 * 8 + (l & 7);
            for (int r = 0; r < 3; ++r) {
                float* row = a.dw + ((size_t)(r * 3 + s) * a.Cin + ci) * a.Cout;
                for (int j = 0; j < (a.Cout >> 3); ++j) {
                    float v[8];
                    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + r * NB + j * 8, v);
                    if (rvalid) {       // 16-byte vector reductions: a quarter of the L2 atomic operations
                        red_add_v4(row + j * 8, v[0], v[1], v[2], v[3]);
                        red_add_v4(row + j * 8 + 4, v[4], v[5], v[6], v[7]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<COLS>(tmem);
}

}  // namespace semb

using namespace semb;

extern "C" int64_t semb_pack_weights_tc(const float* w, int32_t R, int32_t S, int32_t Cin, int32_t Cout, int32_t flip,
                                        void* dst, void* stream) {
    if (!((R == 1 && S == 1) || (R == 3 && S == 3)) || Cin <= 0 || Cout <= 0 || Cin % 8 || Cout % 8) {
        set_error("pack_weights_tc: need 1x1 or 3x3 and 8-padded channels (got %dx%d, %d->%d)", R, S, Cin, Cout);
        return SEMB_ESHAPE;
    }
    const int K = flip ? Cout : Cin, N = flip ? Cin : Cout;
    const TcPlan p = tc_plan(K, N, R * S);
    const long long total = (long long)p.nchunks * p.kchunks * R * S * p.KC * p.NC;
    if (!dst) return total * 2;
    if (!w) { set_error("pack_weights_tc: null weights"); return SEMB_ESHAPE; }
    long long blocks = cdivl(total, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(w, R, S, Cin, Cout, flip, p, reinterpret_cast<bf16*>(dst), total);
    const int rc = check_launch("pack_weights_tc");
    return rc ? rc : total * 2;
}

extern "C" int64_t semb_pack_batch_job_size(void) { return (int64_t)sizeof(PackJob); }

extern "C" int semb_pack_batch_prepare(int32_t index, const float* w, void* dst, int32_t R, int32_t S, int32_t Cin, int32_t Cout,
                                       int32_t flip, int32_t block0, void* host_jobs) {
    SEMB_REQUIRE(host_jobs && w && dst && index >= 0, SEMB_ESHAPE, "pack_batch_prepare: null argument");
    SEMB_REQUIRE(((R == 1 && S == 1) || (R == 3 && S == 3)) && Cin > 0 && Cout > 0 && Cin % 8 == 0 && Cout % 8 == 0, SEMB_ESHAPE,
                 "pack_batch_prepare: need 1x1 or 3x3 and 8-padded channels (got %dx%d, %d->%d)", R, S, Cin, Cout);
    PackJob& j = reinterpret_cast<PackJob*>(host_jobs)[index];
    j.w = w; j.dst = reinterpret_cast<bf16*>(dst); j.R = R; j.S = S; j.Cin = Cin; j.Cout = Cout; j.flip = flip;
    const int K = flip ? Cout : Cin, N = flip ? Cin : Cout;
    j.p = tc_plan(K, N, R * S);
    j.total = (long long)j.p.nchunks * j.p.kchunks * R * S * j.p.KC * j.p.NC;
    long long nb = cdivl(j.total, 2048);
    if (nb > 64) nb = 64;
    if (nb < 1) nb = 1;
    j.block0 = block0;
    j.nblocks = (int)nb;
    return block0 + (int)nb;           // first block of the next job
}

extern "C" int semb_pack_weights_tc_batch(const void* device_jobs, int32_t njobs, int32_t total_blocks, void* stream) {
    SEMB_REQUIRE(device_jobs && njobs > 0 && total_blocks > 0, SEMB_ESHAPE, "pack_weights_tc_batch: bad arguments");
    launch_pdl(pack_weights_batch_kernel, dim3(total_blocks), dim3(256), 0, as_stream(stream), reinterpret_cast<const PackJob*>(device_jobs), njobs);
    return check_launch("pack_weights_tc_batch");
}

extern "C" int semb_conv2d_fwd_tc(const semb_conv_geom* g, const semb_tensor* x, const void* w_packed, const float* bias,
                                  const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                                  int32_t accumulate, void* stream) {
    SEMB_REQUIRE(g && x && y && w_packed, SEMB_ESHAPE, "conv_tc: null argument");
    SEMB_REQUIRE(g->dtype == SEMB_BF16, SEMB_ESHAPE, "conv_tc: bf16 storage only");
    SEMB_REQUIRE(g->stride == 1 && ((g->R == 1 && g->S == 1) || (g->R == 3 && g->S == 3)), SEMB_ESHAPE,
                 "conv_tc: stride-1 1x1 / 3x3 only (got %dx%d stride %d)", g->R, g->S, g->stride);
    SEMB_REQUIRE(view_ok(x) && view_ok(y) && x->C == g->Cin && y->C == g->Cout, SEMB_EALIGN, "conv_tc: bad tensor views");
    SEMB_REQUIRE(g->N > 0 && g->OH > 0 && g->OW > 0 && g->pad_t >= 0 && g->pad_l >= 0 && g->pad_t < g->R + TILE_H && g->pad_l < g->S + TILE_W,
                 SEMB_ESHAPE, "conv_tc: bad geometry");
    // zero padding is the TMA's out-of-bounds fill: conv_tma.cu; reflect padding keeps the cp.async producers below
    if (g->pad_mode == SEMB_PAD_ZERO && !getenv("SEMB_TC_NO_TMA"))
        return conv_tma_launch(g, x, w_packed, bias, y, stats, stats_nstride, stats_cstride, accumulate, stream);
    TcArgs a{};
    a.N = g->N; a.H = g->H; a.W = g->W; a.OH = g->OH; a.OW = g->OW; a.Cin = g->Cin; a.Cout = g->Cout;
    a.R = g->R; a.S = g->S; a.pad_t = g->pad_t; a.pad_l = g->pad_l; a.pad_mode = g->pad_mode;
    a.x = reinterpret_cast<const bf16*>(x->ptr); a.x_pitch = x->pitch; a.x_coff = x->coff;
    a.y = reinterpret_cast<bf16*>(y->ptr); a.y_pitch = y->pitch; a.y_coff = y->coff;
    a.wp = reinterpret_cast<const bf16*>(w_packed); a.bias = bias;
    a.stats = reinterpret_cast<double*>(stats); a.stats_nstride = stats_nstride; a.stats_cstride = stats_cstride;
    a.accumulate = accumulate;
    a.tiles_x = cdiv(g->OW, TILE_W); a.tiles_y = cdiv(g->OH, TILE_H);
    a.total_tiles = g->N * a.tiles_x * a.tiles_y;
    a.p = tc_plan(g->Cin, g->Cout, g->R * g->S);
    a.halo_h = TILE_H + g->R - 1; a.halo_w = TILE_W + g->S - 1;
    a.plane_bytes = a.halo_h * a.halo_w * 16 + 16;          // +16: consecutive planes start in different banks
    const int taps = g->R * g->S;
    a.a_bytes = (a.p.KC / 8) * a.plane_bytes;
    a.b_bytes = taps * a.p.KC * a.p.NC * 2;
    a.b_resident = a.p.kchunks == 1;
    a.stage_bytes = a.a_bytes + (a.b_resident ? 0 : a.b_bytes);
    const size_t fixed = (a.b_resident ? a.b_bytes : 0) + (size_t)8 * a.p.NC * sizeof(float);
    a.stages = (size_t)4 * a.stage_bytes + fixed <= 64 * 1024 ? 4 : ((size_t)3 * a.stage_bytes + fixed <= 200 * 1024 ? 3 : 2);
    const size_t smem = (size_t)a.stages * a.stage_bytes + fixed;
    SEMB_REQUIRE(smem <= 220 * 1024, SEMB_EWORKSPACE, "conv_tc: %zu bytes of shared memory needed", smem);
    const int cols = 2 * a.p.NC <= 32 ? 32 : (2 * a.p.NC <= 64 ? 64 : (2 * a.p.NC <= 128 ? 128 : (2 * a.p.NC <= 256 ? 256 : 512)));
    int per_sm = 512 / cols;                                  // TMEM columns
    if ((size_t)per_sm * (smem + 2048) > 220 * 1024) per_sm = (int)(220 * 1024 / (smem + 2048));   // shared memory
    if (per_sm > 3) per_sm = 3;                               // registers: 288 threads x <= 72
    if (const char* env = getenv("SEMB_TC_DEBUG")) a.dbg = atoi(env);
    if (const char* env = getenv("SEMB_TC_PER_SM")) { const int v = atoi(env); if (v >= 1 && v < per_sm) per_sm = v; }
    if (a.p.NC == 32 && per_sm > 2) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    int gx = 148 * per_sm;
    if (gx > a.total_tiles) gx = a.total_tiles;
    dim3 grid(gx, a.p.nchunks);
    cudaError_t e = cudaSuccess;
#define SEMB_TC_LAUNCH3(COLS, ST, NCT, KR)                                                                                \
    e = cudaFuncSetAttribute(conv_tc_kernel<COLS, ST, NCT, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e == cudaSuccess) conv_tc_kernel<COLS, ST, NCT, KR><<<grid, FW_THREADS, smem, as_stream(stream)>>>(a);
#define SEMB_TC_LAUNCH2(COLS, ST, NCT)                                                                                    \
    if (g->R == 3) { SEMB_TC_LAUNCH3(COLS, ST, NCT, 3) } else { SEMB_TC_LAUNCH3(COLS, ST, NCT, 1) }
#define SEMB_TC_LAUNCH(COLS, NCT)                                                                                         \
    if (a.stages == 4) { SEMB_TC_LAUNCH2(COLS, 4, NCT) } else if (a.stages == 3) { SEMB_TC_LAUNCH2(COLS, 3, NCT) } else { SEMB_TC_LAUNCH2(COLS, 2, NCT) }
    switch (cols) {
        case 32: SEMB_TC_LAUNCH(32, 16) break;       // NC == 16
        case 64: SEMB_TC_LAUNCH(64, 32) break;       // NC == 32
        case 128: SEMB_TC_LAUNCH(128, 0) break;
        case 256: SEMB_TC_LAUNCH(256, 0) break;
        default: SEMB_TC_LAUNCH(512, 0) break;
    }
#undef SEMB_TC_LAUNCH
#undef SEMB_TC_LAUNCH2
#undef SEMB_TC_LAUNCH3
    if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEMB_ECUDA; }
    return check_launch("conv_tc");
}

extern "C" int semb_conv2d_fwd_tc_f32(const semb_conv_geom* g, const semb_tensor* x3, const void* w_packed, const float* bias,
                                      const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                                      int32_t accumulate, void* stream) {
    SEMB_REQUIRE(g && x3 && y && w_packed, SEMB_ESHAPE, "conv_tc_f32: null argument");
    SEMB_REQUIRE(g->dtype == SEMB_BF16 && g->pad_mode == SEMB_PAD_ZERO, SEMB_ESHAPE, "conv_tc_f32: geometry must describe the bf16 split operand, zero padding");
    SEMB_REQUIRE(g->stride == 1 && ((g->R == 1 && g->S == 1) || (g->R == 3 && g->S == 3)), SEMB_ESHAPE,
                 "conv_tc_f32: stride-1 1x1 / 3x3 only (got %dx%d stride %d)", g->R, g->S, g->stride);
    SEMB_REQUIRE(view_ok(x3) && x3->C == g->Cin, SEMB_EALIGN, "conv_tc_f32: bad operand view");
    SEMB_REQUIRE(y->ptr && y->C == g->Cout && (y->C % 8) == 0 && (y->pitch % 8) == 0 && (y->coff % 8) == 0 && y->coff + y->C <= y->pitch &&
                 (reinterpret_cast<uintptr_t>(y->ptr) % 32) == 0, SEMB_EALIGN, "conv_tc_f32: bad fp32 output view");
    SEMB_REQUIRE(g->N > 0 && g->OH > 0 && g->OW > 0 && g->pad_t >= 0 && g->pad_l >= 0 && g->pad_t < g->R + TILE_H && g->pad_l < g->S + TILE_W,
                 SEMB_ESHAPE, "conv_tc_f32: bad geometry");
    return conv_tma_launch(g, x3, w_packed, bias, y, stats, stats_nstride, stats_cstride, accumulate, stream, 1);
}

extern "C" int64_t semb_conv2d_wgrad_tc_workspace(const semb_conv_geom* g) {
    if (!g || g->dtype != SEMB_BF16 || g->stride != 1 || g->R != 3 || g->S != 3 || g->pad_mode != SEMB_PAD_ZERO || g->pad_t > 2 || g->pad_l > 2)
        return 0;
    // planar staging pays where the channel-group count makes the 16-byte NHWC box rows the limit (measured on the CycleGAN layers)
    static const int min_c = [] { const char* e = getenv("SEMB_WGRAD_PLANAR_MIN_C"); return e ? atoi(e) : 128; }();
    if (g->Cin < min_c || g->Cout < min_c) return 0;
    return (int64_t)wgrad_tma_workspace_bytes(g);
}

extern "C" int semb_conv2d_wgrad_tc_ws(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy, float* dw, void* workspace,
                                       int64_t workspace_bytes, void* stream) {
    if (!workspace) return semb_conv2d_wgrad_tc(g, x, dy, dw, stream);
    SEMB_REQUIRE(g && x && dy && dw, SEMB_ESHAPE, "wgrad_tc_ws: null argument");
    const int64_t need = semb_conv2d_wgrad_tc_workspace(g);
    SEMB_REQUIRE(need > 0, SEMB_ESHAPE, "wgrad_tc_ws: this geometry has no planar path (query semb_conv2d_wgrad_tc_workspace first)");
    SEMB_REQUIRE(workspace_bytes >= need, SEMB_EWORKSPACE, "wgrad_tc_ws: %lld bytes of workspace needed, %lld given", (long long)need,
                 (long long)workspace_bytes);
    SEMB_REQUIRE(view_ok(x) && view_ok(dy) && x->C == g->Cin && dy->C == g->Cout, SEMB_EALIGN, "wgrad_tc_ws: bad tensor views");
    return wgrad_tma_launch(g, x, dy, dw, workspace, stream);
}

extern "C" int semb_conv2d_wgrad_tc(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy, float* dw,
                                    void* stream) {
    SEMB_REQUIRE(g && x && dy && dw, SEMB_ESHAPE, "wgrad_tc: null argument");
    SEMB_REQUIRE(g->dtype == SEMB_BF16, SEMB_ESHAPE, "wgrad_tc: bf16 storage only");
    SEMB_REQUIRE(g->stride == 1 && ((g->R == 1 && g->S == 1) || (g->R == 3 && g->S == 3)), SEMB_ESHAPE,
                 "wgrad_tc: stride-1 1x1 / 3x3 only (got %dx%d stride %d)", g->R, g->S, g->stride);
    SEMB_REQUIRE(view_ok(x) && view_ok(dy) && x->C == g->Cin && dy->C == g->Cout, SEMB_EALIGN, "wgrad_tc: bad tensor views");
    if (g->R == 3 && g->pad_mode == SEMB_PAD_ZERO && g->pad_t <= 2 && g->pad_l <= 2 && !getenv("SEMB_WGRAD_NO_TMA"))
        return wgrad_tma_launch(g, x, dy, dw, nullptr, stream);
    if (g->R == 3 && g->Cin <= 40 && g->Cout <= 128 && g->pad_t <= 2 && g->pad_l <= 2 && !getenv("SEMB_WGRAD_NO_STACK")) {
        WsArgs w{};
        w.N = g->N; w.H = g->H; w.W = g->W; w.OH = g->OH; w.OW = g->OW; w.Cin = g->Cin; w.Cout = g->Cout;
        w.pad_t = g->pad_t; w.pad_l = g->pad_l; w.pad_mode = g->pad_mode;
        w.x = reinterpret_cast<const bf16*>(x->ptr); w.x_pitch = x->pitch; w.x_coff = x->coff;
        w.dy = reinterpret_cast<const bf16*>(dy->ptr); w.dy_pitch = dy->pitch; w.dy_coff = dy->coff;
        w.dw = dw;
        w.tiles_x = cdiv(g->OW, TILE_W); w.tiles_y = cdiv(g->OH, TILE_H);
        w.total_tiles = g->N * w.tiles_x * w.tiles_y;
        w.p = g->Cin / 8;
        w.NB = (g->Cout + 15) / 16 * 16;
        w.a_bytes = 3 * w.p * WS_PLANE_A;
        w.stage_bytes = w.a_bytes + (w.NB / 8) * WS_PLANE_B;
        // the M = 128 MMA always reads 16 row groups; groups beyond 3p are garbage rows that are never flushed, but the
        // reads must stay inside the allocation
        const int slack = 16 * WS_PLANE_A > w.stage_bytes ? 16 * WS_PLANE_A - w.stage_bytes : 0;
        const int stages = (size_t)3 * w.stage_bytes + slack <= 110 * 1024 ? 3 : 2;
        const size_t smem = (size_t)stages * w.stage_bytes + slack;
        SEMB_REQUIRE(smem <= 220 * 1024, SEMB_EWORKSPACE, "wgrad_tc(stacked): %zu bytes of shared memory needed", smem);
        const int need = 3 * w.NB;
        const int cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
        int per_sm = 512 / cols;
        if ((size_t)per_sm * (smem + 2048) > 220 * 1024) per_sm = (int)(220 * 1024 / (smem + 2048));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        // split-K over CTAs: every CTA ends with 9*Cin*Cout gradient reductions, so few-tile problems use fewer CTAs.
        // minimise tiles/splits * t_tile + splits * t_flush  (t_tile ~ 1 us, ~60 G reduced floats per second)
        const double t_flush_us = 9.0 * g->Cin * g->Cout / 60e3;
        long long splits = (long long)sqrt((double)w.total_tiles / t_flush_us);
        if (splits > 148LL * per_sm) splits = 148LL * per_sm;
        if (splits > w.total_tiles) splits = w.total_tiles;
        if (splits < 1) splits = 1;
        w.splits = (int)splits;
        cudaError_t e = cudaSuccess;
#define SEMB_WS_LAUNCH(COLS)                                                                                                \
        if (stages == 3) {                                                                                                  \
            e = cudaFuncSetAttribute(wgrad_tc_stacked_kernel<COLS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e == cudaSuccess) wgrad_tc_stacked_kernel<COLS, 3><<<w.splits, WG_THREADS, smem, as_stream(stream)>>>(w);   \
        } else {                                                                                                            \
            e = cudaFuncSetAttribute(wgrad_tc_stacked_kernel<COLS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e == cudaSuccess) wgrad_tc_stacked_kernel<COLS, 2><<<w.splits, WG_THREADS, smem, as_stream(stream)>>>(w);   \
        }
        switch (cols) {
            case 64: SEMB_WS_LAUNCH(64) break;
            case 128: SEMB_WS_LAUNCH(128) break;
            case 256: SEMB_WS_LAUNCH(256) break;
            default: SEMB_WS_LAUNCH(512) break;
        }
#undef SEMB_WS_LAUNCH
        if (e != cudaSuccess) { set_error("wgrad_tc(stacked): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEMB_ECUDA; }
        return check_launch("wgrad_tc_stacked");
    }
    WgArgs a{};
    a.N = g->N; a.H = g->H; a.W = g->W; a.OH = g->OH; a.OW = g->OW; a.Cin = g->Cin; a.Cout = g->Cout;
    a.R = g->R; a.S = g->S; a.pad_t = g->pad_t; a.pad_l = g->pad_l; a.pad_mode = g->pad_mode;
    a.x = reinterpret_cast<const bf16*>(x->ptr); a.x_pitch = x->pitch; a.x_coff = x->coff;
    a.dy = reinterpret_cast<const bf16*>(dy->ptr); a.dy_pitch = dy->pitch; a.dy_coff = dy->coff;
    a.dw = dw;
    a.tiles_x = cdiv(g->OW, TILE_W); a.tiles_y = cdiv(g->OH, TILE_H);
    a.total_tiles = g->N * a.tiles_x * a.tiles_y;
    a.halo_h = TILE_H + g->R - 1; a.halo_w = TILE_W + g->S - 1;
    a.plane_a = TILE_H * TILE_W * 16 + 16;
    a.plane_b = a.halo_h * a.halo_w * 16 + 16;
    a.p = wg_plan(g->Cin, g->Cout, g->R * g->S, a.total_tiles, a.plane_a, a.plane_b);
    const size_t smem = (size_t)a.p.smem_bytes;
    SEMB_REQUIRE(smem <= 220 * 1024, SEMB_EWORKSPACE, "wgrad_tc: %zu bytes of shared memory needed", smem);
    dim3 grid(a.p.splits, a.p.mtiles * a.p.tapgroups * a.p.nchunks);
    cudaError_t e = cudaSuccess;
#define SEMB_WG_LAUNCH(COLS)                                                                                          \
    if (a.p.stages == 3) {                                                                                            \
        e = cudaFuncSetAttribute(wgrad_tc_kernel<COLS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) e = launch_pdl(wgrad_tc_kernel<COLS, 3>, grid, dim3(WG_THREADS), smem, as_stream(stream), a); \
    } else {                                                                                                          \
        e = cudaFuncSetAttribute(wgrad_tc_kernel<COLS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) e = launch_pdl(wgrad_tc_kernel<COLS, 2>, grid, dim3(WG_THREADS), smem, as_stream(stream), a); \
    }
    switch (a.p.cols) {
        case 32: SEMB_WG_LAUNCH(32) break;
        case 64: SEMB_WG_LAUNCH(64) break;
        case 128: SEMB_WG_LAUNCH(128) break;
        case 256: SEMB_WG_LAUNCH(256) break;
        default: SEMB_WG_LAUNCH(512) break;
    }
#undef SEMB_WG_LAUNCH
    if (e != cudaSuccess) { set_error("wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEMB_ECUDA; }
    return check_launch("wgrad_tc");
}
